# 8-GPU runs of the multi-view workloads with the overlapped count exchange (and the serial form beside it)
set -x
mkdir -p gpurun_out
N=${N:-8}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611"
$TR bench.py --gpus $N --steps 10 --warmup 3 --workload pubmed_3v --docs 125000 --no-e2e > gpurun_out/ovl${N}_pubmed.log 2>&1
$TR bench.py --gpus $N --steps 10 --warmup 3 --workload acm_2v --docs 125000 --no-e2e > gpurun_out/ovl${N}_acm1m.log 2>&1
$TR bench.py --gpus $N --steps 10 --warmup 3 --workload acm_2v --docs 125000 --no-e2e --no-overlap > gpurun_out/ser${N}_acm1m.log 2>&1
$TR bench.py --gpus $N --steps 10 --warmup 3 --workload pubmed_3v --docs 125000 --no-e2e --reserve-sms 12 > gpurun_out/ovl${N}_pubmed_r12.log 2>&1
$TR bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/def${N}_lda.log 2>&1
for f in ovl${N}_pubmed ovl${N}_acm1m ser${N}_acm1m ovl${N}_pubmed_r12 def${N}_lda; do tail -1 gpurun_out/$f.log | cut -c1-400; done
