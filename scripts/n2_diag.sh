# usage: bash scripts/n2_diag.sh [N]   diagnostic sweep: CTA limit of the overlapped count all-reduce, and the step without it
N=${1:-2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
for c in 2 1 4; do
  DBN_BENCH_NCCL_CTAS=$c timeout 150 $TR bench.py --gpus $N --steps 100 --warmup 20 > gpurun_out/n${N}_ctas$c.json 2> gpurun_out/n${N}_ctas$c.err
done
DBN_BENCH_DIAG_NOREDUCE=1 timeout 150 $TR bench.py --gpus $N --steps 100 --warmup 20 > gpurun_out/n${N}_noreduce.json 2> gpurun_out/n${N}_noreduce.err
tail -c 300 gpurun_out/n${N}_ctas*.json; tail -3 gpurun_out/n${N}_ctas2.err
