set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/c3_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/c3_tests.log
tail -4 gpurun_out/c3_tests.log
timeout 400 python bench.py > gpurun_out/c3_bench_default.json 2> gpurun_out/c3_bench_default.err; cut -c1-300 gpurun_out/c3_bench_default.json
MMPL_ATTN_HALF=0 timeout 400 python bench.py > gpurun_out/c3_bench_whole.json 2> gpurun_out/c3_bench_whole.err; cut -c1-300 gpurun_out/c3_bench_whole.json
# launch list of one forward (L_kv = 18720 and 32760)
for c in 3 6; do
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/c3_launches_chunk$c.csv python tools/profile_forward.py --chunk $c --forwards 2 > gpurun_out/c3_prof$c.log 2>&1
python tools/launch_summary.py gpurun_out/c3_launches_chunk$c.csv > gpurun_out/c3_launches_chunk$c.txt; head -8 gpurun_out/c3_launches_chunk$c.txt
done
# full capture: attention at L_kv 32760 / 18720 / 512 vs cuDNN
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"flash_attn|sdpa|combine" -o gpurun_out/c3_attn_cmp python tools/attn_vs_cudnn.py 4680x32760x12 4680x18720x12 4680x512x12 > gpurun_out/c3_ncu_attn.log 2>&1
python tools/ncu_summary.py gpurun_out/c3_attn_cmp.ncu-rep > gpurun_out/c3_attn_cmp.txt 2>&1; grep -c "##" gpurun_out/c3_attn_cmp.txt
ls -la gpurun_out | head -40
