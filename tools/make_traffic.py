"""profiles/r02_traffic.json: DRAM bytes (dram__bytes_read.sum + dram__bytes_write.sum, `ncu --set full`) per op call of
each kernel class of the bench step (6 tiles of 512x512 per GPU), from the two committed captures:
  profiles/r02_ncu_full_vit_t6_3layers.csv   ViT kernels at the bench batch (6 tiles, M = 18 912 rows), 3-block ViT
  profiles/r02_ncu_full_jbu_t1.csv           JBU / head kernels for one tile: its 16-crop chunk is exactly a chunk of the
                                             bench step; the image-level kernels see 1/6 of the bench canvas (scaled x6)
bench.py reports the value next to each class's roofline line as `traffic` (bytes per op call, averaged over the class)."""
import csv
import json
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
T = 6


def rows(path):
    r = list(csv.reader(open(os.path.join(ROOT, path))))
    h = r[0]
    ik, ir, iw, it = h.index('Kernel Name'), h.index('dram__bytes_read.sum'), h.index('dram__bytes_write.sum'), h.index('gpu__time_duration.sum')
    u = r[1]
    sc = {'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}
    out = []
    for x in r[2:]:
        out.append((x[ik], float(x[ir]) * sc[u[ir]] + float(x[iw]) * sc[u[iw]], float(x[it])))
    return out


vit = rows('profiles/r02_ncu_full_vit_t6_3layers.csv')
jbu = rows('profiles/r02_ncu_full_jbu_t1.csv')
res = {}

# ---- ViT classes: per-launch averages over the launches of one standard block (the capture holds 2 standard blocks)
gemm_big = [b for n, b, t in vit if 'gemm_bf16_tcgen05_kernel<256' in n and t > 25]       # QKV, out-proj, fc1, fc2, proj, patch-embed
res['gemm'] = sum(gemm_big) / len(gemm_big)
res['layernorm'] = sum(b for n, b, t in vit if 'layernorm_reg_kernel<__nv_bfloat16>' in n) / sum(1 for n, b, t in vit if 'layernorm_reg_kernel<__nv_bfloat16>' in n)
att = [b for n, b, t in vit if 'attention' in n]
res['attention'] = sum(att) / len(att)
res['simmap'] = next(b for n, b, t in vit if 'simmap' in n)

# ---- JBU classes: the capture order is image-level stage 2, image-level stage 3, then the chunk's four stages
img = jbu[1:11]            # after outlier_apply: [proj, range, fixup, tables, composite] x 2
chunk = jbu[11:]
def cls(rs, key):
    return [b for n, b, t in rs if key in n]
per_step = {
    'jbu_range_kernel': (T * sum(cls(img, 'range_kernel_mma')) + T * sum(cls(chunk, 'range_kernel_mma')), 2 + T * 4),
    'jbu_kernel_fixup': (T * sum(cls(img, 'kernel_fixup')) + T * sum(cls(chunk, 'kernel_fixup')), 2 + T * 4),
    'jbu_apply': (T * sum(cls(img, 'fz_composite')) + T * (sum(cls(chunk, 'fz_composite')) + sum(cls(chunk, 'jbu_apply_fused'))), 2 + T * 4),
    'basis_logits': (T * sum(cls(chunk, 'basis_logits')), T),
    'accum_argmax': (T * sum(cls(chunk, 'accum_argmax')), 1),
}
for k, (tot, calls) in per_step.items():
    res[k] = tot / calls
res = {k: round(v) for k, v in res.items()}
res['_note'] = ('bytes per op call, averaged over the calls of the class in one bench step (6 tiles / GPU); '
                'source: ncu --set full captures profiles/r02_ncu_full_*.csv (tools/make_traffic.py)')
json.dump(res, open(os.path.join(ROOT, 'profiles', 'r02_traffic.json'), 'w'), indent=1)
print(json.dumps(res, indent=1))
