#!/bin/bash
# Profiling session on the GPU box (run through gpurun from the repo root): launch list of one bench step, --set full rows of
# every kernel class, and source-level stall samples of the three kernels the step spends most non-GEMM time in.
# ncu reports stay in /tmp on the box; only CSV exports go to gpurun_out/.
set -u
OUT=gpurun_out
TAG=${1:-r02p}
mkdir -p $OUT
NCU="ncu --profile-from-start off --clock-control none"
python tools/profile_step.py --tiles 6 > $OUT/prof_plain.log 2>&1 || { echo "plain run failed"; tail -5 $OUT/prof_plain.log; exit 1; }
$NCU --metrics gpu__time_duration.sum --csv --log-file $OUT/${TAG}_launches_step_t6.csv python tools/profile_step.py --tiles 6 > $OUT/prof_ncu1.log 2>&1; echo "launch list rc=$?"
python tools/profile_step.py --tiles 6 --layers 3 > $OUT/prof_plainA.log 2>&1 && \
$NCU --set full -k 'regex:gemm_bf16|layernorm|attention|simmap|patchify|embed|outlier|cls_debias' -f -o /tmp/${TAG}_full_vit \
    python tools/profile_step.py --tiles 6 --layers 3 > $OUT/prof_ncuA.log 2>&1; echo "full vit rc=$?"
ncu -i /tmp/${TAG}_full_vit.ncu-rep --page raw --csv > /tmp/${TAG}_vit_raw.csv 2>/dev/null && python tools/ncu_trim.py /tmp/${TAG}_vit_raw.csv $OUT/${TAG}_ncu_full_vit_t6_3layers.csv
python tools/profile_step.py --tiles 1 --layers 3 > $OUT/prof_plainB.log 2>&1 && \
$NCU --set full -k 'regex:range|fixup|composite|apply|basis|accum|tables|iou' -f -o /tmp/${TAG}_full_jbu \
    python tools/profile_step.py --tiles 1 --layers 3 > $OUT/prof_ncuB.log 2>&1; echo "full jbu rc=$?"
ncu -i /tmp/${TAG}_full_jbu.ncu-rep --page raw --csv > /tmp/${TAG}_jbu_raw.csv 2>/dev/null && python tools/ncu_trim.py /tmp/${TAG}_jbu_raw.csv $OUT/${TAG}_ncu_full_jbu_t1.csv
if [ "${SKIP_SRC:-0}" != "1" ]; then
  # source-level samples (SASS + line info): one launch each, bench batch
  src() {   # name regex skip
    $NCU --section SourceCounters --section WarpStateStats --section SchedulerStats --import-source on -k "regex:$2" --launch-skip $3 -c 1 -f -o /tmp/${TAG}_src_$1 \
        python tools/profile_step.py --tiles 6 --layers 3 > $OUT/prof_src_$1.log 2>&1
    ncu -i /tmp/${TAG}_src_$1.ncu-rep --page source --csv > $OUT/${TAG}_src_$1.csv 2>/dev/null
    echo "src $1 rc=$? $(wc -c < $OUT/${TAG}_src_$1.csv) bytes"
  }
  src attn_tc 'attention_tc_kernel' 0
  src attn_exp 'attention_tc_exp_kernel' 0
  src range 'range_kernel_mma<5, 128, 0>' 1
  src apply 'jbu_apply_fused_kernel' 3
fi
ls -la $OUT | tail -14
du -sh $OUT
