"""One training step of the dual-NCDE models the way the reference harness runs it.

Reference: experiments/common.py:_train_loop (:378-614) / _train_loop_forecasting (:205-375):
pred = model(times, coeffs, final_index, slope); loss = task loss + 0.03 * sum ||theta_g||
(:47-58, regularisation over func_g only); loss.backward(); optimizer.step().  Data-parallel runs
add one flat-gradient all-reduce between backward and the optimizer step (ancde_b200/dp.py).
"""
import torch

import os

from . import heads, metamodel, solver, synthetic
from .dp import FlatGradBucket


_pos_weight = {}
ASYNC_WGRAD = os.environ.get('ANCDE_ASYNC_WGRAD', '1') not in ('', '0')     # weight-gradient kernels on the library's side stream


def task_loss(cfg, pred, y):
    if cfg['kind'] == 'ancde_forecasting':
        return torch.nn.functional.mse_loss(pred, y)                          # common.py:766
    if cfg['out'] == 1:                                                       # Sepsis: BCE, pos_weight 10
        pw = _pos_weight.get(pred.device)
        if pw is None:      # created once: torch.tensor(..., device=cuda) is a synchronising pageable copy
            pw = _pos_weight[pred.device] = torch.tensor(10.0, device=pred.device)
        return torch.nn.functional.binary_cross_entropy_with_logits(pred.squeeze(-1), y, pos_weight=pw)
    return torch.nn.functional.cross_entropy(pred, y)


class DualNcdeTrainer:
    """Model + Adam + flat gradient bucket for one of `synthetic.CONFIGS`."""

    def __init__(self, workload, device, seed=2021, lr=1e-4, adjoint=False, file='train_h_prime'):
        self.cfg = synthetic.CONFIGS[workload]
        self.workload = workload
        self.device = torch.device(device)
        self.adjoint = adjoint
        torch.manual_seed(seed)                       # same initial weights on every rank
        self.model, self.func_g, self.func_f = metamodel.model_from_config(self.cfg, file=file)
        self.model = self.model.to(self.device)
        self.bucket = FlatGradBucket(self.model.parameters())
        fused = self.device.type == 'cuda'
        # capturable: the step counter lives on the device, so a whole step can be captured in a CUDA graph
        # Adam over the ONE flat parameter buffer (every parameter is a view of it, same lr / weight decay for all, like
        # common.py:690,784): one elementwise kernel instead of a multi-tensor launch over ~40 small tensors
        self.flat_parameter = torch.nn.Parameter(self.bucket.flat_param, requires_grad=True)
        self.flat_parameter.grad = self.bucket.flat
        self.opt = torch.optim.Adam([self.flat_parameter], lr=lr, weight_decay=1e-4, fused=fused, capturable=fused)
        self.reg_params = [p for p in self.func_g.parameters() if p.requires_grad]
        self.reg_scaling = 0.03
        self._reg = self.bucket.segments(self.reg_params) if self.reg_params else None
        # parameter sizes in flat-buffer order (the order `seg` numbers them in); computed here because bincount
        # synchronises with the host, which a CUDA-graph capture of step() would not survive
        self._reg_lengths = torch.bincount(self._reg[2], minlength=len(self.reg_params)) if self._reg else None
        # fused tail (read-out + linear + loss + their gradients in one launch) and regulariser (two launches): GPU only
        self.fused = fused
        self.regulariser = heads.WeightRegulariser(self.bucket, self.reg_params, self.reg_scaling) if (fused and self.reg_params) else None
        core = self.model.model if self.cfg['static'] else self.model
        self.readout_linear = core.linear
        if self.cfg['kind'] == 'ancde_forecasting':
            self.tail_kind = heads.TAIL_MSE
        else:
            self.tail_kind = heads.TAIL_BCE if self.cfg['out'] == 1 else heads.TAIL_CE

    def _coeffs(self, batch):
        if 'coeffs' in batch:
            return batch['coeffs']
        # raw paths: the spline coefficients of the batch are built on the GPU inside the step (one kernel launch,
        # datasets_common.coefficients_on_the_fly) instead of being read from a 4x larger precomputed cache
        from .datasets_common import coefficients_on_the_fly
        return coefficients_on_the_fly(batch['times'], batch['X'])

    def forward(self, batch, slope=1.0):
        cfg = self.cfg
        coeffs = self._coeffs(batch)
        kw = dict(stream=cfg['stream'], adjoint=self.adjoint)
        if cfg['static']:
            return self.model(batch['times'], tuple(coeffs) + (batch['static'],), batch['final_index'], slope, **kw)
        return self.model(batch['times'], tuple(coeffs), batch['final_index'], slope, **kw)

    def hidden(self, batch, slope=1.0):
        """The top NCDE's states at every knot, (L, B, H), with the autograd graph of both solves behind them."""
        cfg = self.cfg
        kw = dict(stream=cfg['stream'], adjoint=self.adjoint)
        coeffs = self._coeffs(batch)
        if cfg['static']:
            return self.model.hidden(batch['times'], tuple(coeffs) + (batch['static'],), batch['final_index'], slope, **kw)
        return self.model.hidden(batch['times'], tuple(coeffs), batch['final_index'], slope, **kw)

    def step(self, batch, slope=1.0):
        """forward + loss + backward + gradient all-reduce + Adam. Returns the (detached) loss.

        On the GPU the tail of the step is fused: read-out, final linear layer, task loss and the gradients of all three are
        ONE launch (heads.readout_loss), the backward pass continues from the cotangent of the top solve's states, and the
        weight regulariser (value and gradient) is two launches on the flat buffers."""
        if self.fused:
            return self.step_fused(batch, slope)
        return self.step_eager(batch, slope)

    def step_fused(self, batch, slope=1.0):
        self.bucket.zero()
        z_t = self.hidden(batch, slope)
        loss, _, grad_z = heads.readout_loss(z_t, batch['final_index'], self.readout_linear, batch['y'], self.tail_kind,
                                             output_time=self.cfg['output_time'], pos_weight=10.0)
        with solver.accumulate_into_param_grads(async_wgrad=ASYNC_WGRAD):      # the solver's weight gradients go straight into the flat bucket
            z_t.backward(grad_z)
        if self.regulariser is not None:
            self.regulariser.add(loss)
        self.bucket.all_reduce_mean()
        self.opt.step()
        return loss

    def step_eager(self, batch, slope=1.0):
        """The same step with the tail as plain torch ops (the reference's formulation; CPU tests and cross-checks)."""
        self.bucket.zero()
        pred = self.forward(batch, slope)
        loss = task_loss(self.cfg, pred, batch['y'])
        with solver.accumulate_into_param_grads():      # the solver's weight gradients go straight into the flat bucket
            loss.backward()
        loss = loss.detach() + self.add_weight_regularisation()
        self.bucket.all_reduce_mean()
        self.opt.step()
        return loss

    def add_weight_regularisation(self):
        """The reference's `loss + 0.03 * sum_p ||p||_2` over func_g (common.py:47-58), value AND gradient, computed on
        the flat parameter / gradient buffers: d||p|| = p / ||p|| (0 for a zero parameter, like torch.norm's backward) is
        added to the gradients after loss.backward().  ~10 kernels instead of ~9 per parameter through autograd."""
        if self._reg is None:
            return 0.0
        a, b, seg = self._reg
        w = self.bucket.flat_param[a:b]
        # one segmented sum (index_add_ on ten bins serialises on atomics: 0.13 ms per step at the Sepsis size)
        sq = torch.segment_reduce(w * w, 'sum', lengths=self._reg_lengths, unsafe=True)
        norms = sq.sqrt()
        coef = torch.where(norms > 0, self.reg_scaling / norms, torch.zeros_like(norms))
        self.bucket.flat[a:b].addcmul_(w, coef[seg])
        return self.reg_scaling * norms.sum()

    def capture(self, batch, slope=1.0, warmup=2):
        """Captures one whole training step on `batch` (forward, loss, backward through the solver, gradient all-reduce,
        Adam) in a CUDA graph: the ~230 launches of a step -- 10 of them the library's kernels, the rest small torch
        kernels of the heads, the loss and the optimizer -- replay without per-launch host overhead.  `batch`'s tensors
        are the graph's static inputs (copy new data into them before a replay); returns a GraphedStep.
        The step has no host synchronisation (cached grid metadata, device-side spline lookup), which capture requires."""
        side = torch.cuda.Stream(self.device)
        side.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(side):
            for _ in range(warmup):                   # allocator warm-up, cached host scalars, lazy CUDA initialisation
                self.step(batch, slope)
        torch.cuda.current_stream(self.device).wait_stream(side)
        torch.cuda.synchronize(self.device)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            loss = self.step(batch, slope)
        return GraphedStep(graph, batch, loss)


class GraphedStep:
    """A captured training step: replay() runs it on whatever the static batch tensors hold now."""

    def __init__(self, graph, batch, loss):
        self.graph, self.batch, self.loss = graph, batch, loss

    def load(self, host_batch):
        """Asynchronous host -> device copy of a (pinned) batch into the static input tensors; returns the bytes copied."""
        if 'times' in host_batch:
            # the captured kernels hold the grid (t0, t_end, step, knot count) of the captured `times` as launch parameters
            raise ValueError("a captured step integrates on the `times` it was captured with; capture a new step for another grid")
        n = 0
        for k, v in host_batch.items():
            dst = self.batch[k]
            if k == 'coeffs':
                for d, c in zip(dst, v):
                    d.copy_(c, non_blocking=True)
                    n += c.numel() * c.element_size()
            else:
                dst.copy_(v, non_blocking=True)
                n += v.numel() * v.element_size()
        return n

    def replay(self):
        self.graph.replay()
        return self.loss
