"""profiles/r02_traffic.json: DRAM bytes (dram__bytes_read.sum + dram__bytes_write.sum, `ncu --set full`) per op call of
each kernel class of the bench step (6 tiles of 512x512 per GPU), from the two committed captures:
  profiles/r02_ncu_full_vit_t6_3layers_v2.csv   ViT kernels at the bench batch (6 tiles, M = 18 912 rows), 3-block ViT
  profiles/r02_ncu_full_jbu_t1_v2.csv           JBU / head kernels for ONE tile (16 crops): every class's bytes scale with
                                                the pixel count, so a bench step (6 tiles) moves 6 x the captured bytes
Kernels are matched by name, the number of op calls per bench step is what bench.py times (`launches_timed` / steps).
bench.py reports the value next to each class's roofline line as `traffic` (bytes per op call, averaged over the class)."""
import csv
import json
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
T = 6
# op calls per bench step (bench.py wraps these ops; 12 ViT blocks, 2 JBU chunks of 48 crops x 4 stages + 2 image-level stages)
CALLS = dict(gemm=54, layernorm=24, attention=12, simmap=1, jbu_range_kernel=10, jbu_kernel_fixup=10, jbu_apply=10,
             basis_logits=2, accum_argmax=1)


def rows(path):
    r = list(csv.reader(open(os.path.join(ROOT, path))))
    h, u = r[0], r[1]
    ik, ir, iw, it = h.index('Kernel Name'), h.index('dram__bytes_read.sum'), h.index('dram__bytes_write.sum'), h.index('gpu__time_duration.sum')
    sc = {'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}
    return [(x[ik], float(x[ir]) * sc[u[ir]] + float(x[iw]) * sc[u[iw]], float(x[it])) for x in r[2:]]


vit = rows('profiles/r02_ncu_full_vit_t6_3layers_v2.csv')
jbu = rows('profiles/r02_ncu_full_jbu_t1_v2.csv')


def mean(xs):
    xs = list(xs)
    return sum(xs) / len(xs)


res = {}
# ---- ViT classes: per-launch averages at the bench batch
res['gemm'] = mean(b for n, b, t in vit if 'gemm_bf16_tcgen05_kernel<256' in n and t > 25)      # QKV, out-proj, fc1, fc2, patch-embed
res['layernorm'] = mean(b for n, b, t in vit if 'layernorm_reg_kernel<__nv_bfloat16>' in n)
res['attention'] = mean(b for n, b, t in vit if 'attention' in n)
res['simmap'] = sum(b for n, b, t in vit if 'simmap_split' in n or 'kernel<128, 4, 0, 0, 3>' in n)   # split + block-diagonal GEMM
# ---- JBU classes: bytes of the one-tile capture x 6, divided by the op calls of a bench step
jsum = lambda *keys: sum(b for n, b, t in jbu if any(k in n for k in keys))
res['jbu_range_kernel'] = T * jsum('range_kernel_mma') / CALLS['jbu_range_kernel']
res['jbu_kernel_fixup'] = T * jsum('kernel_fixup') / CALLS['jbu_kernel_fixup']
res['jbu_apply'] = T * jsum('fz_composite', 'jbu_apply_fused', 'fz_tables') / CALLS['jbu_apply']
res['basis_logits'] = T * jsum('basis_logits') / CALLS['basis_logits']
res['accum_argmax'] = T * jsum('accum_argmax') / CALLS['accum_argmax']
res = {k: round(v) for k, v in res.items()}
res['_note'] = ('bytes per op call, averaged over the calls of the class in one bench step (6 tiles / GPU); '
                'source: ncu --set full captures profiles/r02_ncu_full_*_v2.csv (tools/make_traffic.py)')
json.dump(res, open(os.path.join(ROOT, 'profiles', 'r02_traffic.json'), 'w'), indent=1)
print(json.dumps(res, indent=1))
