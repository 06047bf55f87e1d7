import sys,json
for l in open(sys.argv[1]):
    if l.startswith('{'):
        d=json.loads(l); print(round(d["value"],2), round(d["ms_per_step"],3), 'e2e', round(d["e2e"]["value"],2)); print(d["roofline"]["per_tile_ms_by_kernel"]); print(d["roofline"].get("gemm_by_shape"))
        for r in d.get('rooflines',[]): print(r['kernel'], r['bound'], round(r['frac'],3), round(r['share_of_step'],3))
