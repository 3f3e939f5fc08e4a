set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 tools/run_segment_parallel.py --model 14B --segments 8 --sampling-steps 4 --cfg-pair > gpurun_out/c4_seg8_cfgpair.json 2> gpurun_out/c4_seg8_cfgpair.err
echo rc=$?; cut -c1-600 gpurun_out/c4_seg8_cfgpair.json; tail -5 gpurun_out/c4_seg8_cfgpair.err
