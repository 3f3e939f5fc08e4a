timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > gpurun_out/pytest_final.txt; cat gpurun_out/pytest_final.txt
python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; tail -c 300 gpurun_out/bench_final.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_final_reference.json 2> gpurun_out/bench_final_reference.err
python -c "
import json
d=json.load(open('gpurun_out/bench_final.json')); print(d['value'], d['e2e']['value'], d['clocks'], d['roofline']['achieved'], d['roofline']['frac']); print(d['breakdown']['sites']); print(d['cpu_baseline'])
print(open('gpurun_out/bench_final_reference.json').read()[:300])
"
timeout 600 ncu --set full --clock-control none --profile-from-start off -k regex:"qk_norm_rope|ln_kernel|rmsnorm|combine" -c 8 -f -o gpurun_out/pw_final python tools/profile_forward.py --chunk 3 --forwards 2 > gpurun_out/pw_final.log 2>&1; tail -2 gpurun_out/pw_final.log
