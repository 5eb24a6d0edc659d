#!/bin/bash
# Round-2 evidence batch (one gpurun call; keeps gpurun_out/ under the 64 MiB pull limit: every .ncu-rep is exported to
# CSV pages on the box and deleted). Produces what profiles/README.md lists under "Round 2":
#   isolated timings of the non-GEMM kernels and of every GEMM flavour, the decoder profile,
#   ncu --set full of each non-GEMM kernel, of the ten codebook-search launches of one encode and of the three
#   deferred-LayerNorm GEMM flavours, and (argument "launches") the launch list of scales 8-9 of one timed step of the
#   default bench command.
O=gpurun_out/ev; mkdir -p $O
for k in ln sample embed quant next qkv; do python tools/kernels_one.py $k 10; done > $O/kernels_alone.txt 2>&1
python tools/gemm_one.py > $O/gemm_one_d30.txt 2>&1
python tools/gemm_one.py 16 125 680 > $O/gemm_one_d16.txt 2>&1
python tools/attn_time.py > $O/attn_time.txt 2>&1
python tools/decoder_experiment.py > $O/decoder.txt 2>&1
NCU="ncu --set full --clock-control none"
export_rep() { # name
  ncu -i $O/ncu_$1.ncu-rep --page details --csv > $O/ncu_$1.details.csv 2>/dev/null
  ncu -i $O/ncu_$1.ncu-rep --page raw --csv > $O/ncu_$1.raw.csv 2>/dev/null
  rm -f $O/ncu_$1.ncu-rep
}
cap() { # name kernel-regex skip count kind
  $NCU -k regex:$2 -s $3 -c $4 -o $O/ncu_$1 -f python tools/kernels_one.py $5 1 > $O/ncu_$1.log 2>&1
  export_rep $1
}
cap ln ln_modulate 3 1 ln
cap sample sample_kernel 3 1 sample
cap embed embed_kernel 3 1 embed
cap quant_search quant_search 30 10 quant      # the ten scales of one B=64 encode
cap quant_step 'quant_kernel' 30 3 quant
cap next quant_kernel 30 3 next
$NCU --kernel-name-base demangled -k 'regex:gemm_bf16_kernel' -c 6 -o $O/ncu_gemm_lnf -f python tools/gemm_prof_lnf.py > $O/ncu_gemm_lnf.log 2>&1
export_rep gemm_lnf
if [ "$1" = "launches" ]; then
VAR_B200_PROFILE_STEP=1 VAR_B200_PROFILE_FROM_SCALE=8 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
  --log-file $O/launches_default_last2scales.csv python bench.py --no-cpu --no-secondary --steps 2 --warmup 3 > $O/launches_bench.log 2>&1
fi
du -sh gpurun_out; ls -la $O
