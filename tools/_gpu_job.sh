python tools/finesweep.py c4,c5s,c2 > gpurun_out/r4e_finesweep.txt 2>&1
SB_FINE_SMEM_KB=32 ncu --set full --clock-control none --import-source on -k regex:sb_eval_fine_kernel -s 2 -c 1 -f -o gpurun_out/prof_r4e_c4_fine python tools/evalprof.py c4 hist 128 > gpurun_out/r4e_ncu.log 2>&1
python tools/ncu_metrics.py gpurun_out/prof_r4e_c4_fine.ncu-rep > gpurun_out/r4e_ncu_full_metrics_c4_fine.txt 2>&1
python tools/ncu_lines.py gpurun_out/prof_r4e_c4_fine.ncu-rep sbayes_b200/csrc/libsbayes_b200.so sb_eval_fine_kernelILi7E > gpurun_out/r4e_ncu_source_lines_c4_fine.txt 2>&1
rm -f gpurun_out/prof_r4e_c4_fine.ncu-rep
cat gpurun_out/r4e_finesweep.txt; tail -3 gpurun_out/r4e_ncu.log; head -30 gpurun_out/r4e_ncu_source_lines_c4_fine.txt
