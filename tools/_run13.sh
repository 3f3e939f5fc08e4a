timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q -k "flash_attn" 2>&1 | tail -6 > gpurun_out/pytest_r13.txt; cat gpurun_out/pytest_r13.txt
export ATTN_NO_FA2=1 ATTN_CUDNN=0 ATTN_SPLITS=0
export ATTN_SHAPES=4680x9360x12,4680x14040x12,4680x18720x12,4680x32760x12,10920x14040x40
for rep in 1 2; do
echo "== inline"; python tools/bench_kernels.py --what attn 2>&1 | grep "^attn"
echo "== kernel"; MMPL_ATTN_MERGE=kernel python tools/bench_kernels.py --what attn 2>&1 | grep "^attn"
done > gpurun_out/attn_merge.txt 2>&1; cat gpurun_out/attn_merge.txt
python bench.py --no-cpu-baseline > gpurun_out/bench_r13.json 2> gpurun_out/bench_r13.err
MMPL_ATTN_MERGE=kernel python bench.py --no-cpu-baseline > gpurun_out/bench_r13_k.json 2> gpurun_out/bench_r13_k.err
python -c "
import json
for f in ['bench_r13','bench_r13_k']:
    d=json.load(open('gpurun_out/%s.json'%f)); print(f, d['value'], d['e2e']['value'], d['clocks'], d['breakdown']['self_attn'])
"
