#!/bin/bash
# time the FIR-down kernels under ncu (gpu__time_duration + dram bytes), fp32 and bf16, streaming (s) vs tiled (t)
mkdir -p gpurun_out
for d in fp32 bf16; do for mode in ${FIR_MODES:-s}; do
  USE_B200_FIR_DOWN=$mode USE_B200_OVERLAP=1 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:gn_fir_down -c 2 --csv --log-file gpurun_out/fir_${d}_${mode}.csv python tools/ncu_target.py --batch 4 --dtype $d > /dev/null 2>&1
done; done
python - <<'PY'
import csv,glob
for f in sorted(glob.glob('gpurun_out/fir_*_?.csv')):
    lines=[l for l in open(f) if not l.startswith('==')]
    per={}
    for r in csv.DictReader(lines):
        per.setdefault(r['ID'],{})[r['Metric Name']]=float(r['Metric Value'].replace(',','')); per[r['ID']]['k']=r['Kernel Name'][:40]; per[r['ID']]['u_'+r['Metric Name']]=r['Metric Unit']
    for i,m in per.items():
        t=m['gpu__time_duration.sum']; tu=m['u_gpu__time_duration.sum']
        t_us = t/1e3 if tu in('ns','nsecond') else t
        b=m['dram__bytes_read.sum']+m['dram__bytes_write.sum']
        scale={'byte':1,'Kbyte':1e3,'Mbyte':1e6,'Gbyte':1e9}[m['u_dram__bytes_read.sum']]
        print(f.split('/')[-1], i, m['k'], '%.1f us'%t_us, '%.0f MB'%(b*scale/1e6), '%.2f TB/s'%(b*scale/t_us/1e6))
PY
