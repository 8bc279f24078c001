mkdir -p gpurun_out
nvidia-smi -L | tee gpurun_out/r2a_gpus.log
LSK_TEST_EXPERIMENTAL=1 timeout 900 python -m pytest tests/test_gpu_zz_experimental.py -q -s -x --timeout 300 2>&1 | tail -40 | tee gpurun_out/r2a_experimental.log
for v in 0 1; do LSK_ATTN_PUSH=$v timeout 300 python tools/profile_round.py llama2-7b 20 400 2>&1 | tail -2 | tee -a gpurun_out/r2a_ab.log; done
LSK_LMHEAD_TC=1 timeout 300 python tools/profile_round.py llama2-7b 20 400 2>&1 | tail -2 | tee -a gpurun_out/r2a_ab.log
LSK_PROFILE_CLASSES=1 timeout 300 python tools/profile_round.py llama2-7b 20 400 2>&1 | tail -3 | tee -a gpurun_out/r2a_ab.log
