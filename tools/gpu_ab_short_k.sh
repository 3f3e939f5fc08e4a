#!/bin/bash
# One B200 call: parity of the short-K GEMM candidates (tile 192 / 1128 / 1192, gemm_tcgen05.cu), their A/B against the
# 128 x 128 tiles in isolation and inside the cfg2 step, then the whole GPU suite. Every leg has its own timeout, a
# candidate that fails its tests is left out of the timing legs. Output: gpurun_out/${TAG}_*.
TAG=${1:-r2d1}
OUT=gpurun_out
mkdir -p $OUT
export PYTHONUNBUFFERED=1
python -c "import torch; torch.zeros(1, device='cuda'); print(torch.cuda.get_device_name(0))"  # pages the image in
ok=""
for v in 192 1128 1192; do
  timeout 75 python -m pytest tests/test_kernels_gpu.py -q -m gpu -x -k "(gemm_bias or gemm_epilogues) and -${v}]" \
    > $OUT/${TAG}_gemm_tests_$v.log 2>&1
  rc=$?
  echo "pytest rc=$rc" >> $OUT/${TAG}_gemm_tests_$v.log
  [ $rc -eq 0 ] && ok="$ok $v"
done
echo "candidates that passed:$ok" > $OUT/${TAG}_summary.txt
if [ -n "$ok" ]; then
  tiles=$(echo 128 $ok 448 | tr ' ' ',')
  GEMM_SHAPES=4680x1536x1536 GEMM_TILES=$tiles timeout 60 python tools/bench_kernels.py --what gemm \
    > $OUT/${TAG}_gemm_ab.txt 2>&1
  for v in 128 $ok; do
    MMPL_GEMM_SHORT_K=$v timeout 75 python bench.py --no-chain --no-cpu-baseline --steps 3 \
      > $OUT/${TAG}_bench_sk$v.json 2> $OUT/${TAG}_bench_sk$v.err
  done
fi
deselect=""
for v in 192 1128 1192; do
  case " $ok " in *" $v "*) ;; *) deselect="$deselect and not -${v}]" ;; esac
done
timeout 240 python -m pytest tests -q -m gpu -x -k "not zzz $deselect" > $OUT/${TAG}_tests.log 2>&1
echo "pytest rc=$?" >> $OUT/${TAG}_tests.log
tail -3 $OUT/${TAG}_tests.log
cat $OUT/${TAG}_summary.txt
[ -f $OUT/${TAG}_gemm_ab.txt ] && cat $OUT/${TAG}_gemm_ab.txt
