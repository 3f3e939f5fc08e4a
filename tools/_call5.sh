set -x
N=$1; SEG=$2; EXTRA=$3
mkdir -p gpurun_out
if [ "$N" = "1" ]; then
timeout 300 python tools/run_segment_parallel.py --model 14B --segments $SEG --sampling-steps 4 > gpurun_out/c5_seg_n$N.json 2> gpurun_out/c5_seg_n$N.err
else
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tools/run_segment_parallel.py --model 14B --segments $SEG --sampling-steps 4 $EXTRA > gpurun_out/c5_seg_n$N.json 2> gpurun_out/c5_seg_n$N.err
fi
echo rc=$?; cut -c1-400 gpurun_out/c5_seg_n$N.json; tail -3 gpurun_out/c5_seg_n$N.err
