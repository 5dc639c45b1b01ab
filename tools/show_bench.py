"""Dev helper: one-screen summary of a bench.py JSON line."""
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print('samples/s %.0f  ms/step %.2f  e2e %.0f  step_frac %s  launches %s  graph %s' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['step_frac'], d.get('gpu_launches'), d['config'].get('cuda_graph')))
tot = 0.0
for k in d['kernels']:
    tot += k['avg_ms'] * k['launches'] / d.get('kernel_profile_steps', d['steps'])
    print('  %-26s %8.4f ms x%-3d %s' % (k['name'], k['avg_ms'], k['launches'], k.get('frac_of_fp32_peak')))
print('  kernel ms per step %.2f' % tot)
