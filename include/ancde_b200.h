/*
 * ancde_b200.h -- C ABI of libancde_b200.so (hand-written sm_100a kernels for the
 * ANCDE dual-NCDE hot path).  Plain pointers and sizes only; no torch types.
 *
 * Ownership: the caller allocates every input, output and workspace buffer (device
 * memory, contiguous); the library never allocates or frees device memory, never
 * synchronises the host, and launches on the stream it is given.  Every entry point
 * returns 0 on success or a negative ANCDE_E* code; the message of the last error of
 * the calling thread is available through ancde_last_error().
 *
 * Each entry point cites the reference interface it replaces (paths relative to the
 * reference repository sheoyon-jhin/ANCDE).
 */
#ifndef ANCDE_B200_H
#define ANCDE_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ANCDE_OK            0
#define ANCDE_EINVAL       -1   /* bad argument (shape, null pointer, unsupported size) */
#define ANCDE_ECUDA        -2   /* a CUDA runtime call failed */
#define ANCDE_EUNSUPPORTED -3   /* valid request the kernels do not cover (no fallback exists) */

#define ANCDE_MAX_LAYERS 8      /* hidden (ReLU) layers in a vector-field MLP */

/* ABI version; bumped when a struct below changes (the engine value ANCDE_ENGINE_LANE of round 2 changed none).  7: head / tail / regulariser entry points.  6: ancde_ncde_desc gained `grid`.  5: ancde_ncde_desc gained single_stage, grad_k_out (ABI 5) and
 * async_wgrad (ABI 4, with ancde_ncde_join). */
int ancde_abi_version(void);

/* Copies the calling thread's last error message into buf (NUL terminated). Returns its length. */
int ancde_last_error(char* buf, size_t len);

/* ---------------------------------------------------------------------------------------------
 * Natural cubic spline coefficients with missing values.
 * Replaces controldiffeq.natural_cubic_spline_coeffs (controldiffeq/interpolate.py:159-226),
 * i.e. _natural_cubic_spline_coeffs_without_missing_values (:7-53), the per-scalar-path NaN
 * routine (:78-153) and misc.tridiagonal_solve (controldiffeq/misc.py:12-66).
 *
 *   times : (L)        knot times, strictly increasing
 *   X     : (N, L, C)  values, NaN = missing
 *   a, b, two_c, three_d : (N, L-1, C) outputs
 * times_were_f32 (f64 entry only): the reference keeps `times` in float32 while X is float64
 * (Stock); reciprocals of time differences are then rounded through float32 like it does.
 * ------------------------------------------------------------------------------------------- */
int ancde_spline_coeffs_f32(const float* times, const float* X, float* a, float* b, float* two_c,
                            float* three_d, int64_t N, int L, int C, void* stream);
int ancde_spline_coeffs_f64(const double* times, const double* X, double* a, double* b, double* two_c,
                            double* three_d, int64_t N, int L, int C, int times_were_f32, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Fused neural-CDE solve: z' = tanh(MLP(z)).view(Hd, C) @ dU/dt, fixed-grid RK4 (3/8 rule).
 *
 * One descriptor covers the three reference operators:
 *   mode ANCDE_MODE_PLAIN   controldiffeq.cdeint  + VectorField          (cdeint_module.py:9-32,141-151)
 *   mode ANCDE_MODE_BOTTOM  controldiffeq.ancde_bottom + VectorField_stack (:34-72,168-243); also
 *                           emits h' = every stage derivative in call order (the reference's .npy file)
 *   mode ANCDE_MODE_TOP     controldiffeq.ancde   + AttentiveVectorField  (:74-138,247-254):
 *                           dU = dX*a + a(1-a)*dX*h'[call counter], a = attention[floor(t)-1]
 * The MLP is FinalTanh_ff6 / FinalTanh (experiments/models/vector_fields.py:185-250): n_hidden
 * ReLU layers then a linear layer to Hd*C followed by tanh.  The solver arithmetic is
 * torchdiffeq==0.0.1's FixedGridODESolver with rk4_alt_step_func.
 * ------------------------------------------------------------------------------------------- */
#define ANCDE_MODE_PLAIN  0
#define ANCDE_MODE_BOTTOM 1
#define ANCDE_MODE_TOP    2

/* Two implementations of the same operator (identical inputs/outputs; the workspace sizes differ, so query them with
 * the same `engine` value that the solve will use):
 *   ANCDE_ENGINE_STREAMED  (default) one CTA per tile of 8 samples; the output-layer weights are streamed from L2 through
 *                          a TMA ring every stage.  Forward: FP32 FMA.  Backward and weight gradients: tensor cores
 *                          (mma.sync, FP16 hi/lo split, three products per tile = FP32-grade accuracy).
 *   (value 2 was the round-1 mma.sync cluster engine; it was superseded by ANCDE_ENGINE_UMMA and removed)            */
#define ANCDE_ENGINE_STREAMED 1
/*   ANCDE_ENGINE_UMMA      forward on the tcgen05 tensor cores (accumulators in tensor memory): a cluster of 2/4/8 CTAs per
 *                          16/32/64 samples keeps the output layer resident in shared memory as FP16 hi/lo operands; three
 *                          products per k step = FP32-grade accuracy.  Saves the activations in the streamed engine's
 *                          layout, so ancde_ncde_backward (streamed kernels) works on its output.  Chosen by default
 *                          (engine = 0) when the shape is supported and the output layer has >= 50 000 weights; state and hidden widths <= 63.   */
#define ANCDE_ENGINE_UMMA     3
/*   ANCDE_ENGINE_LANE      small networks (round 2): ONE SMALL CTA integrates one sample, forward and backward -- a thread per
 *                          output-layer column, the RK4 state in registers, FP32 FMA; the latency regime of small batches (UEA:
 *                          32 samples x 724 dependent stages) and, with 4-7 CTAs resident per SM, large batches of small networks.
 *                          C <= 32 with pad4(C) a power of two, state and hidden widths <= 64, Hd * pad4(C) <= 256.  Saves in tiles
 *                          of 8 samples (the tcgen05 weight-gradient kernels' layout; samples_per_cta 0 or 8) or of 1 sample
 *                          (samples_per_cta 1).  Chosen by default (engine = 0) for such shapes when B < 240, and for B >= 240
 *                          when the output layer has <= 8192 weights.                                                            */
#define ANCDE_ENGINE_LANE     4

typedef struct ancde_ncde_desc {
    /* ---- shapes ---- */
    int32_t B;                 /* samples */
    int32_t L;                 /* spline knots; coefficient tensors are (B, L-1, C) */
    int32_t C;                 /* control channels */
    int32_t Hd;                /* state size */
    int32_t n_hidden;          /* hidden layers, 1..ANCDE_MAX_LAYERS */
    int32_t width[ANCDE_MAX_LAYERS];   /* output width of hidden layer l */
    int32_t mode;              /* ANCDE_MODE_* */
    int32_t att_C;             /* TOP: 1 (timewise) or C (elementwise) */
    int32_t att_L;             /* TOP: leading size of attention (knots) */
    int32_t n_hp;              /* TOP: leading size of h_prime (4 * bottom steps) */
    /* ---- the fixed grid: t_k = t_start + k*step (last clipped to t_end) ---- */
    float   t_start, t_end, step;
    int32_t n_steps;           /* grid intervals */
    int32_t n_out;             /* requested output times (t_out[0] == t_start) */
    /* ---- inputs (device pointers, float32) ---- */
    const float* times;        /* (L) */
    const float* t_out;        /* (n_out) increasing */
    const float* coeff_b;      /* (B, L-1, C) */
    const float* coeff_c2;     /* (B, L-1, C) */
    const float* coeff_d3;     /* (B, L-1, C) */
    const float* z0;           /* (B, Hd) */
    const float* W[ANCDE_MAX_LAYERS + 1];     /* torch Linear weights, (out, in) row-major; [n_hidden] = output layer (Hd*C, width[n_hidden-1]) */
    const float* bias[ANCDE_MAX_LAYERS + 1];
    const float* attention;    /* TOP: (att_L, B, att_C) */
    const float* h_prime;      /* TOP: (n_hp, B, C) */
    /* ---- outputs ---- */
    float* z_out;              /* (n_out, B, Hd) */
    float* k_out;              /* BOTTOM (or NULL): (4*n_steps, B, Hd) stage derivatives in call order */
    /* ---- workspace ---- */
    float* wpack;              /* packed/transposed weights, ancde_ncde_wpack_floats() floats */
    float* save_act;           /* NULL for inference; else ancde_ncde_save_act_floats() floats */
    float* save_G;             /* NULL for inference; else ancde_ncde_save_G_floats() floats */
    /* ---- backward only ---- */
    const float* grad_z_out;   /* (n_out, B, Hd) cotangents of z_out */
    float* grad_z0;            /* (B, Hd) written */
    float* grad_W[ANCDE_MAX_LAYERS + 1];      /* accumulated into (caller zeroes), same layout as W */
    float* grad_bias[ANCDE_MAX_LAYERS + 1];
    float* grad_attention;     /* TOP, or NULL (adjoint-mode gradient flow): (att_L, B, att_C) accumulated into */
    int32_t samples_per_cta;   /* samples per tile of the saved layout: 0 = let the library choose; streamed engine 8, 4, 2 or 1; lane engine 8 or 1 */
    int32_t engine;            /* 0 = library default; ANCDE_ENGINE_STREAMED, _UMMA or _LANE (see above) */
    int32_t single_stage;      /* 1 = evaluate only the vector field at (t_start, z0): k_out (1, B, Hd) = f(z0) dX/dt(t_start), the
                                * building block of adaptive solvers driven from the host (dopri5).  n_steps = n_out = 1; streamed
                                * engine.  Backward: grad_k_out (B, Hd) is the cotangent of k_out, grad_z_out is not read. */
    const float* grad_k_out;   /* single_stage backward only */
    int32_t async_wgrad;       /* backward only: 1 = launch the weight-gradient kernels on the library's side stream (they only
                                * depend on the recurrent kernel of the same call) so that they overlap whatever the caller
                                * enqueues next; the caller must call ancde_ncde_join() on its stream before it reads the
                                * gradients, and keep every buffer of the descriptor alive until then.  0 = same stream. */
    const float* grid;         /* NULL: the arithmetic grid above.  Else n_steps + 1 increasing grid points (device), grid[0] = t_start,
                                * grid[n_steps] = t_end: torchdiffeq's `options['grid_constructor']`, or rk4 without a step size
                                * (the grid is `t` itself; FixedGridODESolver.__init__).  `step` is not read then. */
} ancde_ncde_desc;

/* Workspace sizes (in floats) for a descriptor whose shape fields are filled in. */
int64_t ancde_ncde_wpack_floats(const ancde_ncde_desc* d);
int64_t ancde_ncde_save_act_floats(const ancde_ncde_desc* d);
int64_t ancde_ncde_save_G_floats(const ancde_ncde_desc* d);

/* The engine a descriptor resolves to (ANCDE_ENGINE_*; shape fields and `engine` filled in), or -1 when the requested
 * engine does not support the shape.  With engine = 0 this is the library's automatic choice. */
int ancde_ncde_engine(const ancde_ncde_desc* d);

/* Forward solve.  Writes z_out (and k_out, save_act, save_G when non-NULL). */
int ancde_ncde_forward(const ancde_ncde_desc* d, void* stream);

/* Backprop through the solver (discrete adjoint of the RK4 recurrence), using save_act / save_G
 * written by ancde_ncde_forward with the same descriptor.  Overwrites the stage-input slots of save_act
 * with the stage cotangents (consumed by the weight-gradient kernel launched from the same call). */
int ancde_ncde_backward(const ancde_ncde_desc* d, void* stream);

/* Threading: one host thread per device at a time.  The side stream and its fork / done events are shared per device, so
 * two host threads driving backward calls with async_wgrad = 1 on the SAME device would interleave their event records
 * (creation and the pending flag are mutex-protected, the event order is not); use one thread per device, or
 * async_wgrad = 0.  Calls on different devices are independent.
 *
 * Makes `stream` wait for every weight-gradient kernel that backward calls with async_wgrad = 1 put on the side stream
 * of the current device (an event wait: no host synchronisation, capturable in a CUDA graph).  No-op when none is pending. */
int ancde_ncde_join(void* stream);

/* ---------------------------------------------------------------------------------------------
 * The small layers either side of the two solves, fused.
 *
 * HEAD -- experiments/models/metamodel.py:327-348 (ANCDE) / :650-671 (ANCDE_forecasting):
 *   attention = act(time_attention(raw))  (timewise, one value per knot and sample)  or  act(raw)  (elementwise),
 *   act = sigmoid (soft, :332-333) | round(Hardsigmoid(slope * x)) (hard, slope_check, :335-338) | round(sigmoid(x)) (:340-343),
 *   the rounding with the straight-through gradient of RoundFunctionST (:148-161), Hardsigmoid = (hardtanh + 1) / 2 (:139-146);
 *   y0 = feature_extractor(X(t0) * attention[0])  (:345-348).
 * ------------------------------------------------------------------------------------------- */
#define ANCDE_HEAD_SOFT         0
#define ANCDE_HEAD_HARD_STE     1
#define ANCDE_HEAD_HARD_SIGMOID 2

typedef struct ancde_head_desc {
    int32_t L, B, C, H;         /* knots, samples, channels (= bottom state size), top state size */
    int32_t timewise;           /* 1: attention (L, B, 1) = act(raw . ta_w + ta_b); 0: attention (L, B, C) = act(raw) */
    int32_t kind;               /* ANCDE_HEAD_* */
    float   slope;              /* ANCDE_HEAD_HARD_STE only */
    const float* raw;           /* (L, B, C) bottom NCDE states */
    const float* ta_w;          /* (C)  time_attention.weight, timewise only */
    const float* ta_b;          /* (1)  time_attention.bias */
    const float* x0;            /* X(t0): sample b, channel c at x0[b * x0_stride + c] (e.g. the `a` coefficient of the first piece) */
    int64_t x0_stride;
    const float* fe_w;          /* (H, C) feature_extractor.weight */
    const float* fe_b;          /* (H) */
    float* attention;           /* out (L, B, 1 | C) */
    float* y0;                  /* out (B, H) */
    /* backward only */
    const float* grad_attention;/* (L, B, 1 | C) or NULL (adjoint-mode gradient flow: only y0 carries a cotangent) */
    const float* grad_y0;       /* (B, H) */
    float* grad_raw;            /* (L, B, C) written */
    float* grad_ta_w;           /* accumulated into (caller zeroes); timewise only */
    float* grad_ta_b;
    float* grad_fe_w;
    float* grad_fe_b;
} ancde_head_desc;

int ancde_head_forward(const ancde_head_desc* d, void* stream);
/* Reads `attention` as written by the forward (row 0) and recomputes the pre-activations from `raw`. */
int ancde_head_backward(const ancde_head_desc* d, void* stream);

/* TAIL -- read-out, final linear layer, task loss and the gradients of all three in ONE launch:
 *   classification (metamodel.py:360-371): z = z_t[final_index[b], b, :], pred = linear(z);
 *   forecasting    (metamodel.py:683-696): pred[b, q] = linear(z_t[n_out - T + q, b, :]), q < T = output_time;
 *   loss = BCEWithLogits(pos_weight) (experiments/common.py:677-678) | cross entropy (:680) | MSE (:766), mean reduced. */
#define ANCDE_TAIL_BCE 0
#define ANCDE_TAIL_CE  1
#define ANCDE_TAIL_MSE 2

typedef struct ancde_tail_desc {
    int32_t n_out, B, H, O;     /* z_t is (n_out, B, H); the linear layer maps H -> O */
    int32_t kind;               /* ANCDE_TAIL_* */
    int32_t T;                  /* ANCDE_TAIL_MSE: output_time */
    float   pos_weight;         /* ANCDE_TAIL_BCE */
    float   loss_scale;         /* cotangent of the loss (1 for loss.backward()) */
    const float* z_t;
    const int64_t* final_index; /* (B), classification */
    const float* W;             /* (O, H) */
    const float* b;             /* (O) */
    const void* target;         /* BCE: float (B); CE: int64 (B); MSE: float (B, T, O) */
    float* pred;                /* out (B, O) | (B, T, O) */
    float* loss;                /* out (1): written */
    float* grad_z_t;            /* out (n_out, B, H), written completely (zero where nothing is read out), or NULL */
    float* grad_W;              /* accumulated into (caller zeroes) */
    float* grad_b;
} ancde_tail_desc;

int ancde_tail_forward_backward(const ancde_tail_desc* d, void* stream);

/* REG -- experiments/common.py:47-58: loss + scaling * sum_p ||p||_2 over parameters that are segments
 * [seg_begin[s], seg_begin[s+1]) of one flat buffer: adds scaling * p / ||p|| to flat_grad and scaling * sum ||p|| to *loss
 * (loss may be NULL).  sumsq: nseg floats of scratch. */
int ancde_weight_regulariser(const float* flat_param, float* flat_grad, const int32_t* seg_begin, int32_t nseg, float* sumsq,
                             float scaling, float* loss, void* stream);

/* Number of kernel launches the last forward / backward call of this thread issued. */
int ancde_last_launch_count(void);

/* Per-kernel timing for the roofline report: while enabled, every kernel launch of this library is
 * bracketed by CUDA events on its stream.  After the caller has synchronised the stream,
 * ancde_profile_get(i, ...) returns the name and duration (ms) of the i-th launch since enabling. */
int ancde_profile_enable(int on);
int ancde_profile_count(void);
int ancde_profile_get(int i, char* name, size_t len, float* ms);

/* Development / test aid: 1 = compute the output-layer weight gradient with the mma.sync kernel (ncde_wgrad_mma.cu)
 * instead of the tcgen05 kernel (ncde_wgrad_umma.cu), which is the default for tiles of 8 samples.  Returns the old value. */
int ancde_debug_force_wgrad_mma_sync(int on);

/* Development / test aid: which recurrent backward kernel ancde_ncde_backward runs.  engine 0 = automatic (the cluster /
 * tcgen05 kernel of ncde_bwd4.cu for large output layers, the streamed kernel of ncde_bwd.cu otherwise), 1 = streamed,
 * 4 = cluster (ANCDE_EUNSUPPORTED for shapes it cannot run), 5 = cluster wherever it supports the shape, else streamed;
 * cluster_size 0 = automatic, else 2, 4 or 8.  Returns the old engine value.  ancde_debug_bwd4_launches counts the
 * launches of the cluster kernel since the library was loaded. */
int ancde_debug_force_bwd_engine(int engine, int cluster_size);
int ancde_debug_bwd4_launches(void);

/* Development aid: 64 int64 counters in device memory that kernels built with -DANCDE_PHASE_TIMING
 * fill with per-phase clock64() totals (normal builds never touch the buffer). NULL disables. */
int ancde_debug_set_buffer(long long* dev_counters);

/* FP32 FMA throughput microbenchmark (roofline denominator): launches a kernel of
 * dependent-free FFMA chains; returns the FLOP count it executes in *flops. */
int ancde_fp32_peak_kernel(float* scratch, int iters, double* flops, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* ANCDE_B200_H */
