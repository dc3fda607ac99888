mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_model.py tests/test_gpu_dropins.py -q --no-header -rf -x -p no:cacheprovider -k "attention or train or loss or grad or dropout or evaluator or samplenet or decoding" > gpurun_out/pytest_train.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/pytest_train.log | cut -c1-300
timeout 300 python bench.py --config cfg4 --no-cpu-baseline --no-parity > gpurun_out/bench_cfg4.log 2> gpurun_out/bench_cfg4.err; echo "rc=$?"; tail -1 gpurun_out/bench_cfg4.log | cut -c1-200
