TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
$TR bench.py --gpus 2 --steps 100 --warmup 20 > gpurun_out/n2_default.json 2> gpurun_out/n2_default.err
DBN_BENCH_DIAG_NOREDUCE=1 $TR bench.py --gpus 2 --steps 100 --warmup 20 > gpurun_out/n2_noreduce.json 2> gpurun_out/n2_noreduce.err
NCCL_MAX_NCHANNELS=1 $TR bench.py --gpus 2 --steps 100 --warmup 20 > gpurun_out/n2_1chan.json 2> gpurun_out/n2_1chan.err
CUDA_VISIBLE_DEVICES=1 python bench.py --steps 100 --warmup 20 --no-cpu-baseline > gpurun_out/n1_dev1.json 2> gpurun_out/n1_dev1.err
tail -c 600 gpurun_out/n2_*.json gpurun_out/n1_dev1.json
