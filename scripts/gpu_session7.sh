#!/bin/bash
mkdir -p gpurun_out
MODLE_B200_REGISTER_PATH=direct timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sectors_op_red.sum,smsp__inst_executed_op_global_red.sum --clock-control none -k regex:"k_register" -c 8 --csv --log-file gpurun_out/register_launches2.csv python scripts/bench_register.py --reps 0 > gpurun_out/register_ncu2.log 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/register_launches2.csv')) if len(r)>10]
hdr=rows[0]
ik=hdr.index('Kernel Name'); im=hdr.index('Metric Name'); iv=hdr.index('Metric Value'); iu=hdr.index('Metric Unit'); iid=hdr.index('ID')
from collections import OrderedDict
d=OrderedDict()
for r in rows[1:]:
    d.setdefault((r[iid],r[ik][:24]),{})[r[im]]=r[iv]
for k,v in d.items(): print(k, v)
PY
tail -3 gpurun_out/register_ncu2.log | cut -c1-200
