# 2-GPU weak-scaling step under different NCCL CTA budgets / DP variants (profiling)
one() { python -c "
import json,sys; d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); print(sys.argv[1], round(d['value']), round(d['ms_per_step'],3))" $1; }
run() { name=$1; shift; env "$@" timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $((29540 + RANDOM % 200)) bench.py --gpus 2 --steps 30 --warmup 5 --skip-cpu --skip-extras > gpurun_out/r02ae_$name.json 2>> gpurun_out/r02ae.err; one gpurun_out/r02ae_$name.json; }
CUDA_VISIBLE_DEVICES=0 timeout 300 python bench.py --steps 30 --warmup 5 --skip-cpu --skip-extras > gpurun_out/r02ae_n1.json 2> gpurun_out/r02ae.err; one gpurun_out/r02ae_n1.json
run default MOEGAN_X=0
run ctas2 NCCL_MAX_CTAS=2
run ctas4 NCCL_MAX_CTAS=4
run ctas8 NCCL_MAX_CTAS=8
run localbal MOEGAN_BENCH_LOCAL_BALANCE=1
run localbal_ctas4 MOEGAN_BENCH_LOCAL_BALANCE=1 NCCL_MAX_CTAS=4
run single_ctas32 MOEGAN_DP_MODE=single NCCL_MIN_CTAS=32
