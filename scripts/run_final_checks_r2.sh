# round-2 end-state checks on one B200: the GPU suite, the 21-case 3e7-photon parity sweep, the reference-count anchors
timeout 1200 python -m pytest tests -x -q -m gpu > gpurun_out/r2_gpu_tests_final.txt 2>&1; tail -3 gpurun_out/r2_gpu_tests_final.txt
timeout 900 python scripts/highstat.py 30000000 > gpurun_out/r2_parity_sweep_3e7.txt 2>&1; tail -4 gpurun_out/r2_parity_sweep_3e7.txt
timeout 900 python scripts/anchors.py 20000 > gpurun_out/r2_anchors.txt 2>&1; tail -4 gpurun_out/r2_anchors.txt
