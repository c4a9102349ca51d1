# full GPU suite on the current build, then ncu --set full of the two persistent loop kernels (T = 8, B = 64: same CTAs as the headline)
python -m pytest tests -m gpu -q --timeout 1500 -x > gpurun_out/r3e_gpu_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r3e_gpu_tests.log
tail -6 gpurun_out/r3e_gpu_tests.log
for L in 1 2; do
  PTK_FWD_LOOP=$L timeout 120 python tools/prof_step.py 8 64 1 > gpurun_out/r3e_plain$L.log 2>&1 &&
  PTK_FWD_LOOP=$L timeout 600 ncu --set full --clock-control none --import-source on -k regex:"int_loop" -c 1 -o gpurun_out/r3e_loop$L python tools/prof_step.py 8 64 1 > gpurun_out/r3e_ncu$L.log 2>&1
  tail -2 gpurun_out/r3e_ncu$L.log
done
ls -la gpurun_out/r3e_*
