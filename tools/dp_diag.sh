# per-GPU single-process step (GPU 0, GPU 1, both at once) next to the 2-GPU data-parallel step on the same box
one() { python -c "
import json,sys; d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); print(sys.argv[1], round(d['value']), round(d['ms_per_step'],3), d['clocks'])" $1; }
B="python bench.py --steps 30 --warmup 5 --skip-cpu --skip-extras"
CUDA_VISIBLE_DEVICES=0 timeout 300 $B > gpurun_out/r02ad_gpu0_alone.json 2> gpurun_out/r02ad.err; one gpurun_out/r02ad_gpu0_alone.json
CUDA_VISIBLE_DEVICES=1 timeout 300 $B > gpurun_out/r02ad_gpu1_alone.json 2>> gpurun_out/r02ad.err; one gpurun_out/r02ad_gpu1_alone.json
CUDA_VISIBLE_DEVICES=0 timeout 300 $B > gpurun_out/r02ad_gpu0_conc.json 2>> gpurun_out/r02ad.err &
P0=$!
CUDA_VISIBLE_DEVICES=1 timeout 300 $B > gpurun_out/r02ad_gpu1_conc.json 2>> gpurun_out/r02ad.err &
P1=$!
wait $P0; wait $P1
one gpurun_out/r02ad_gpu0_conc.json; one gpurun_out/r02ad_gpu1_conc.json
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 30 --warmup 5 --skip-cpu --skip-extras > gpurun_out/r02ad_dp2.json 2>> gpurun_out/r02ad.err; one gpurun_out/r02ad_dp2.json
