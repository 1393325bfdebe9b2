set -x
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "async or overlapped or invariants" > gpurun_out/pytest_async.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_async.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611"
$TR bench.py --gpus 2 --steps 10 --warmup 3 --workload pubmed_3v --docs 125000 --no-e2e > gpurun_out/ovl2_pubmed.log 2>&1
$TR bench.py --gpus 2 --steps 10 --warmup 3 --workload pubmed_3v --docs 125000 --no-e2e --no-overlap > gpurun_out/ser2_pubmed.log 2>&1
$TR bench.py --gpus 2 --steps 10 --warmup 3 --workload pubmed_3v --docs 125000 --no-e2e --reserve-sms 16 > gpurun_out/ovl2_pubmed_r16.log 2>&1
tail -3 gpurun_out/pytest_async.log; tail -1 gpurun_out/ovl2_pubmed.log; tail -1 gpurun_out/ser2_pubmed.log; tail -1 gpurun_out/ovl2_pubmed_r16.log
