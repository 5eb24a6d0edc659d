#!/bin/bash
# DRAM bytes per launch of the three deferred-LayerNorm GEMM flavours under different settings (ncu, 4 metrics)
# usage: tools/dram_probe.sh "ENV=VAL ENV2=VAL" ...   (one quoted environment per variant)
for envs in "$@"; do
  env $envs ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active --clock-control none --kernel-name-base demangled -k 'regex:gemm_bf16_kernel' -c 6 --csv python tools/gemm_prof_lnf.py 2>/dev/null | python -c "
import csv,sys
rows=[r for r in csv.DictReader(l for l in sys.stdin if l.startswith('\"'))]
per={}
for r in rows: per.setdefault(r['ID'],{})[r['Metric Name']]=float(r['Metric Value'].replace(',',''))
names=['fc1','fc1','proj','proj','fc2','fc2']
for i,(k,v) in enumerate(sorted(per.items(), key=lambda kv:int(kv[0]))):
    print('$envs |', names[i], 'dram read %.3f write %.3f GB, %.1f us, tensor pipe %.1f%%'%(v['dram__bytes_read.sum']/1e9, v['dram__bytes_write.sum']/1e9, v['gpu__time_duration.sum']/1e3, v['sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active']))
"
done
