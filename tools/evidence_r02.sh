# Final evidence of round 2 (one B200): run under gpurun from the repo root, e.g.
#   gpurun --timeout 1500 -- 'bash tools/evidence_r02.sh r02m'
# Everything lands in gpurun_out/<tag>_*; copy what should be judged into profiles/.
set -x
T=${1:-r02m}
cd ${GRAFT_REPO_ROOT:-.}
O=gpurun_out
timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -15 > $O/${T}_pytest_gpu.log
tail -1 $O/${T}_pytest_gpu.log
# ---- bench lines (never under a profiler)
timeout 400 python bench.py --steps 50 --warmup 5 > $O/${T}_bench_swinir_x4.json 2> $O/${T}_bench.err
for w in hat_x4 dat_x2; do timeout 300 python bench.py --workload $w --steps 20 --warmup 5 > $O/${T}_bench_$w.json 2>> $O/${T}_bench.err; done
timeout 400 python bench.py --impl reference --steps 2 --warmup 1 > $O/${T}_bench_reference.json 2>> $O/${T}_bench.err
timeout 120 python tools/kbench_pair.py > $O/${T}_kbench_pair.log 2>&1
timeout 120 python tools/kbench_conv.py > $O/${T}_kbench_conv.log 2>&1
timeout 150 python tools/tight_bench.py 10 > $O/${T}_tight_bench.log 2>&1
timeout 200 python tools/step_bench.py > $O/${T}_step_bench.log 2>&1
SRK_LIB=$PWD/tpu_superresolution_b200/lib/libsrk_dbg.so timeout 120 python tools/timeline.py > $O/${T}_timeline.log 2>&1
# ---- ncu launch lists of one steady-state forward per family (after the plain run has exited 0)
for w in swinir_x4 hat_x4 dat_x2; do
  python tools/run_forward_any.py $w 3 > $O/plain_$w.log 2>&1 && \
  ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file $O/${T}_${w}_ncu_launches.csv python tools/run_forward_any.py $w 3 > $O/ncu_$w.log 2>&1
  python tools/ncu_launch_summary.py $O/${T}_${w}_ncu_launches.csv > $O/${T}_${w}_ncu_launch_summary.txt
done
# ---- ncu --set full of the top kernels (one launch each)
full() {  # name, kernel regex, workload, skip
  python tools/run_forward_any.py $3 3 > $O/plain_full.log 2>&1 && \
  ncu --set full --clock-control none --import-source on --profile-from-start off -k "regex:$2" -s $4 -c 1 -o $O/${T}_prof_$1 -f python tools/run_forward_any.py $3 3 > $O/ncu_full_$1.log 2>&1
}
full attn swin_attn_kernel swinir_x4 2
full mlp swin_mlp_kernel swinir_x4 2
full conv conv3x3_kernel swinir_x4 1
full winattn winattn_kernel hat_x4 2
full linear token_linear_kernel dat_x2 3
ls -la $O/${T}_*
