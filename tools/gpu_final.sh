set -u
OUT=gpurun_out
timeout 500 python -m pytest tests -m gpu -q -x > $OUT/r02bs_tests.log 2>&1; echo "tests rc=$?"; tail -2 $OUT/r02bs_tests.log
timeout 300 python bench.py > $OUT/r02bs_bench_default.log 2>&1; echo "bench rc=$?"
timeout 200 python bench.py --impl reference --steps 2 --warmup 1 > $OUT/r02bs_bench_reference.log 2>&1; echo "ref rc=$?"
python -c "import __graft_entry__ as g; g.smoke()" > $OUT/r02bs_smoke.log 2>&1; echo "smoke rc=$?"
timeout 100 python tools/profile_step.py --tiles 6 > $OUT/prof_plain.log 2>&1 && timeout 200 ncu --profile-from-start off --clock-control none --metrics gpu__time_duration.sum --csv --log-file $OUT/r02_launches_step_t6_v7.csv python tools/profile_step.py --tiles 6 > $OUT/prof_ncu1.log 2>&1; echo "launch list rc=$?"
python tools/showbench.py $OUT/r02bs_bench_default.log | head -1
