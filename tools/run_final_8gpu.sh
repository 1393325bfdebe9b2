# 8-GPU final lines: the two 1 M-document multi-view shapes, the default (driver) workload, and BASELINE configs[4] in full
set -x
mkdir -p gpurun_out
N=${N:-8}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611"
$TR bench.py --gpus $N --steps 10 --warmup 3 --workload acm_2v --docs 125000 --no-e2e > gpurun_out/v9_${N}_acm1m.log 2>&1
$TR bench.py --gpus $N --steps 10 --warmup 3 --workload pubmed_3v --docs 125000 --no-e2e > gpurun_out/v9_${N}_pubmed1m.log 2>&1
$TR bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/v9_${N}_lda.log 2>&1
timeout 420 $TR bench.py --gpus $N --steps 5 --warmup 3 --workload stress_4v --docs 250000 --no-e2e > gpurun_out/v9_${N}_stress2m.log 2>&1
for f in acm1m pubmed1m lda stress2m; do tail -1 gpurun_out/v9_${N}_$f.log | cut -c1-260; done
