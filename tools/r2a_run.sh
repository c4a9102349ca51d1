set -x
nvidia-smi -L
python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/r2a_gpu_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2a_gpu_tests.log
tail -30 gpurun_out/r2a_gpu_tests.log
python tools/parity_report.py > gpurun_out/r2a_parity_report.txt 2>&1
tail -12 gpurun_out/r2a_parity_report.txt
python bench.py --steps 10 --warmup 3 > gpurun_out/r2a_bench.log 2>&1; tail -c 1500 gpurun_out/r2a_bench.log
for tool in memcheck racecheck synccheck; do
  timeout 600 compute-sanitizer --tool $tool python tools/sanitize_small.py > gpurun_out/r2a_san_${tool}_small.log 2>&1; echo "rc=$?" >> gpurun_out/r2a_san_${tool}_small.log
  tail -5 gpurun_out/r2a_san_${tool}_small.log
done
timeout 600 compute-sanitizer --tool memcheck python tools/sanitize_wide.py > gpurun_out/r2a_san_memcheck_wide.log 2>&1; echo "rc=$?" >> gpurun_out/r2a_san_memcheck_wide.log
tail -5 gpurun_out/r2a_san_memcheck_wide.log
