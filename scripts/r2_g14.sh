mkdir -p gpurun_out
nvidia-smi -L | head -3
timeout 600 python -m pytest tests/test_host_shim.py tests/test_c5_partition.py -m gpu -x -q 2>&1 | tail -3
./tests/cpp/comm_test
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2_g14_c3_n2.json 2> gpurun_out/r2_g14_c3_n2.err; tail -c 2200 gpurun_out/r2_g14_c3_n2.json; echo; tail -5 gpurun_out/r2_g14_c3_n2.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --workload c5 --steps 10 --warmup 3 > gpurun_out/r2_g14_c5_n2.json 2> gpurun_out/r2_g14_c5_n2.err; python -c "
import json; d=json.loads(open('gpurun_out/r2_g14_c5_n2.json').read().strip().splitlines()[-1]); print('c5 n2', d['ms_per_step'], d['e2e']['ms_per_step'])"
