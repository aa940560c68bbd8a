set -x
timeout 600 python -m pytest tests/test_gpu_distributed.py -x -q -m gpu > gpurun_out/r2_dist_tests_n8.txt 2>&1; tail -3 gpurun_out/r2_dist_tests_n8.txt
timeout 300 python bench.py --gpus 1 --steps 20 --warmup 3 --no-cpu-baseline --no-drivers > gpurun_out/r2_scale_n1.json 2> gpurun_out/r2_scale_n1.err
for n in 2 4 8; do timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 20 --warmup 3 > gpurun_out/r2_scale_n$n.json 2> gpurun_out/r2_scale_n$n.err; done
timeout 300 python bench.py --gpus 1 --workload config4 --photons 100000 --steps 5 --warmup 3 > gpurun_out/r2_strong_n1.json 2> gpurun_out/r2_strong_n1.err
for n in 2 4 8; do timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2952$n bench.py --gpus $n --workload config4 --photons 100000 --steps 5 --warmup 3 > gpurun_out/r2_strong_n$n.json 2> gpurun_out/r2_strong_n$n.err; done
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29538 bench.py --gpus 8 --workload config5 --steps 20 --warmup 3 > gpurun_out/r2_config5_n8.json 2> gpurun_out/r2_config5_n8.err
for f in gpurun_out/r2_scale_n*.json gpurun_out/r2_strong_n*.json gpurun_out/r2_config5_n8.json; do echo $f; cut -c1-220 $f; done
