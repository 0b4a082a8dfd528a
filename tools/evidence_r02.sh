set -x
cd $GRAFT_REPO_ROOT
timeout 300 python -m pytest tests/test_validate_tool.py -q -m gpu 2>&1 | tail -3
for w in swinir_x4 hat_x4 dat_x2; do
  python tools/run_forward_any.py $w 3 > gpurun_out/plain_$w.log 2>&1 && \
  ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r02_${w}_ncu_launches.csv python tools/run_forward_any.py $w 3 > gpurun_out/ncu_$w.log 2>&1
  tail -1 gpurun_out/ncu_$w.log
done
python tools/run_forward_any.py swinir_x4 3 > gpurun_out/plain_full.log 2>&1 && \
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:conv3x3_kernel -s 1 -c 1 -o gpurun_out/r02_prof_conv python tools/run_forward_any.py swinir_x4 3 > gpurun_out/ncu_full1.log 2>&1
python tools/run_forward_any.py swinir_x4 3 > gpurun_out/plain_full.log 2>&1 && \
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:swin_attn_kernel -s 2 -c 1 -o gpurun_out/r02_prof_attn python tools/run_forward_any.py swinir_x4 3 > gpurun_out/ncu_full2.log 2>&1
python tools/run_forward_any.py swinir_x4 3 > gpurun_out/plain_full.log 2>&1 && \
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:swin_mlp_kernel -s 2 -c 1 -o gpurun_out/r02_prof_mlp python tools/run_forward_any.py swinir_x4 3 > gpurun_out/ncu_full3.log 2>&1
python tools/run_forward_any.py hat_x4 3 > gpurun_out/plain_full.log 2>&1 && \
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:winattn_kernel -s 2 -c 1 -o gpurun_out/r02_prof_winattn python tools/run_forward_any.py hat_x4 3 > gpurun_out/ncu_full4.log 2>&1
python tools/run_forward_any.py dat_x2 3 > gpurun_out/plain_full.log 2>&1 && \
ncu --set full --clock-control none --profile-from-start off -k "regex:dwconv3x3_rows|dat_mix|channel_gram|channel_apply|token_linear|row_stats|rows_to_f16|cab_gate" -s 8 -c 12 -o gpurun_out/r02_prof_dat_aux python tools/run_forward_any.py dat_x2 3 > gpurun_out/ncu_full5.log 2>&1
ls -la gpurun_out/*.ncu-rep
