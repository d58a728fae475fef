#!/bin/bash
mkdir -p gpurun_out
run() { # name chroms streams
  MODLE_B200_BENCH_CHROMS=$2 timeout 600 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --streams $3 > gpurun_out/s19_$1.json 2> gpurun_out/s19_$1.err
  python -c "
import json
d=json.loads([l for l in open('gpurun_out/s19_$1.json') if l.startswith('{')][-1]); print('$1 chroms=$2 streams=$3', 'value %.1f M/s'%(d['value']/1e6), 'ms %.1f'%d['ms_per_step'], 'avg_launch_ms %.1f'%d['roofline']['avg_launch_ms'], 'cyc/cell-epoch %.0f'%d['cycles_per_cell_epoch'])
"
}
run one_chr1 chr1 1
run one_chr2 chr2 1
run one_chr3 chr3 1
run three_serial chr1,chr2,chr3 1
run three_s3 chr1,chr2,chr3 3
run six_s1 chr1,chr2,chr3,chr4,chr5,chr6 1
run six_s3 chr1,chr2,chr3,chr4,chr5,chr6 3
run six_s6 chr1,chr2,chr3,chr4,chr5,chr6 6
run small_s1 chr17,chr18,chr19,chr20,chr21,chr22 1
run small_s3 chr17,chr18,chr19,chr20,chr21,chr22 3
run small_s6 chr17,chr18,chr19,chr20,chr21,chr22 6
