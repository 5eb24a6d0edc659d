O=gpurun_out/ev3; mkdir -p $O
python __graft_entry__.py smoke 2>&1 | tail -2
ncu --set full --clock-control none --kernel-name-base demangled -k 'regex:gemm_bf16_kernel' -c 6 -o $O/ncu_gemm_lnf -f python tools/gemm_prof_lnf.py > $O/ncu_gemm_lnf.log 2>&1
ncu -i $O/ncu_gemm_lnf.ncu-rep --page details --csv > $O/ncu_gemm_lnf.details.csv
ncu -i $O/ncu_gemm_lnf.ncu-rep --page raw --csv > $O/ncu_gemm_lnf.raw.csv
rm -f $O/ncu_gemm_lnf.ncu-rep
ncu --set full --clock-control none -k regex:quant_search -s 30 -c 10 -o $O/ncu_qs -f python tools/kernels_one.py quant 1 > $O/ncu_qs.log 2>&1
ncu -i $O/ncu_qs.ncu-rep --page details --csv > $O/ncu_qs.details.csv; ncu -i $O/ncu_qs.ncu-rep --page raw --csv > $O/ncu_qs.raw.csv; rm -f $O/ncu_qs.ncu-rep
tools/ab_env.sh "--workload score_d16 --steps 5 --warmup 3" VAR_B200_PDL=1 VAR_B200_PDL=0 VAR_B200_PDL=1 VAR_B200_PDL=0
du -sh gpurun_out
