timeout 400 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k 'regex:pair_kernel<.int.1, .int.1' -s 5 -c 2 -f -o gpurun_out/r02b_pair_c3 python tools/run_plan.py resnet50 1 > gpurun_out/ncu_c3.log 2>&1
tail -3 gpurun_out/ncu_c3.log; ls -la gpurun_out/r02b_pair_c3.ncu-rep
