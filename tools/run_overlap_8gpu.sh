# 8-GPU runs of the multi-view workloads with the overlapped count exchange
set -x
mkdir -p gpurun_out
N=${N:-8}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611"
$TR bench.py --gpus $N --steps 10 --warmup 3 --workload pubmed_3v --docs 125000 --no-e2e > gpurun_out/ovlw${N}_pubmed.log 2>&1
$TR bench.py --gpus $N --steps 10 --warmup 3 --workload acm_2v --docs 125000 --no-e2e > gpurun_out/ovlw${N}_acm1m.log 2>&1
$TR bench.py --gpus $N --steps 10 --warmup 3 --workload acm_2v --docs 125000 --no-e2e --narrow-all > gpurun_out/ovln${N}_acm1m.log 2>&1
for f in ovlw${N}_pubmed ovlw${N}_acm1m ovln${N}_acm1m; do tail -1 gpurun_out/$f.log | cut -c1-300; done
