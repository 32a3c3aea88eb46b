# is the data-parallel overhead the collectives, or the presence of NCCL nodes in the captured graph?  (profiling)
one() { python -c "
import json,sys; d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); print(sys.argv[1], round(d['value']), round(d['ms_per_step'],3))" $1; }
run() { name=$1; shift; env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $((29540 + RANDOM % 200)) bench.py --gpus 2 --steps 30 --warmup 5 --skip-cpu --skip-extras > gpurun_out/r02ag_$name.json 2>> gpurun_out/r02ag.err; one gpurun_out/r02ag_$name.json; }
CUDA_VISIBLE_DEVICES=0 timeout 300 python bench.py --steps 30 --warmup 5 --skip-cpu --skip-extras > gpurun_out/r02ag_n1.json 2> gpurun_out/r02ag.err; one gpurun_out/r02ag_n1.json
run default MOEGAN_X=0
run skip_all MOEGAN_DP_SKIP_COLLECTIVE=1 MOEGAN_BENCH_LOCAL_BALANCE=1
run skip_grads MOEGAN_DP_SKIP_COLLECTIVE=1

