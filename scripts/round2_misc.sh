mkdir -p gpurun_out/round2
timeout 300 python scripts/stream_diag.py 2>&1 | grep -v Warn | tee gpurun_out/round2/stream_diag.log
timeout 400 python scripts/bench_configs.py c1 2>&1 | grep -v Warn | tee gpurun_out/round2/c1_variants.txt
timeout 300 python -m pytest tests/test_parity_gpu.py -q -k "native_maths" 2>&1 | tail -5 | tee gpurun_out/round2/native.log
bash scripts/round2_sanitize.sh 2>&1 | tail -15
