#!/bin/bash
# One GPU-box session: tests, bench lines, reference comparator, ncu evidence.  Run through gpurun from the repo root.
# ncu reports stay in /tmp on the box (a --set full report is > 64 MiB); only CSV exports go to gpurun_out/.
set -u
OUT=gpurun_out
TAG=${1:-r02}
mkdir -p $OUT
timeout 600 python -m pytest tests -m gpu -q -s --maxfail=10 > $OUT/${TAG}_tests.log 2>&1; echo "tests rc=$?"; tail -3 $OUT/${TAG}_tests.log
timeout 300 python bench.py --steps 10 --warmup 3 --tiles 6 > $OUT/${TAG}_bench_t6.log 2>&1; echo "bench6 rc=$?"
CSEG_ATTN_TC=0 timeout 300 python bench.py --steps 10 --warmup 3 --tiles 6 --no-cpu-baseline > $OUT/${TAG}_bench_t6_noattntc.log 2>&1; echo "bench6 attn-mma rc=$?"
CSEG_JBU_SHARE=0 timeout 300 python bench.py --steps 10 --warmup 3 --tiles 6 --no-cpu-baseline > $OUT/${TAG}_bench_t6_noshare.log 2>&1; echo "bench6 noshare rc=$?"
timeout 300 python bench.py --steps 10 --warmup 3 --tiles 4 --no-cpu-baseline > $OUT/${TAG}_bench_t4.log 2>&1; echo "bench4 rc=$?"
timeout 300 python bench.py --steps 5 --warmup 3 --tiles 2 --workload road1024 > $OUT/${TAG}_bench_road.log 2>&1; echo "road rc=$?"
timeout 300 python bench.py --steps 5 --warmup 3 --tiles 2 --workload isaid896 > $OUT/${TAG}_bench_isaid.log 2>&1; echo "isaid rc=$?"
if [ "${SKIP_REF:-0}" != "1" ]; then
  python -m oracle.ref_on_gpu --out $OUT/${TAG}_ref_on_gpu.json > $OUT/${TAG}_ref_on_gpu.log 2>&1; echo "ref rc=$?"
fi
if [ "${SKIP_NCU:-0}" != "1" ]; then
  python tools/profile_step.py --tiles 6 > $OUT/prof_plain.log 2>&1 && \
  ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/${TAG}_launches_step_t6.csv \
      python tools/profile_step.py --tiles 6 > $OUT/prof_ncu1.log 2>&1; echo "ncu launch list rc=$?"
  python tools/profile_step.py --tiles 6 --layers 3 > $OUT/prof_plainA.log 2>&1 && \
  ncu --profile-from-start off --set full --clock-control none -k 'regex:gemm_bf16|layernorm|attention|simmap|patchify' -f -o /tmp/${TAG}_full_vit \
      python tools/profile_step.py --tiles 6 --layers 3 > $OUT/prof_ncuA.log 2>&1; echo "ncu full vit rc=$?"
  ncu -i /tmp/${TAG}_full_vit.ncu-rep --page raw --csv > $OUT/${TAG}_ncu_full_vit_t6.csv 2>/dev/null
  python tools/profile_step.py --tiles 1 --layers 3 > $OUT/prof_plainB.log 2>&1 && \
  ncu --profile-from-start off --set full --clock-control none -k 'regex:range|fixup|composite|apply|basis|accum|tables|iou' -f -o /tmp/${TAG}_full_jbu \
      python tools/profile_step.py --tiles 1 --layers 3 > $OUT/prof_ncuB.log 2>&1; echo "ncu full jbu rc=$?"
  ncu -i /tmp/${TAG}_full_jbu.ncu-rep --page raw --csv > $OUT/${TAG}_ncu_full_jbu_t1.csv 2>/dev/null
fi
ls -la $OUT | tail -20
du -sh $OUT
