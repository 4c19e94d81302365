#!/bin/bash
# Run on the GPU box (under gpurun): launch list of one train step, then per-launch ncu metrics of exactly one step.
#   bash profiles/capture_step.sh <tag>      -> gpurun_out/<tag>_launches.csv, gpurun_out/<tag>_ncu_step_raw.md
set -e
TAG=${1:-r2}
ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python profiles/one_step.py 3 > /dev/null 2>&1
N=$(python - <<PY
import csv
rows = list(csv.DictReader(l for l in open('gpurun_out/${TAG}_launches.csv') if not l.startswith('==')))
idx = [i for i, r in enumerate(rows) if 'adam_kernel' in r['Kernel Name']]
print(idx[-1] - idx[-2])
PY
)
echo "launches per step: $N"
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__throughput.avg.pct_of_peak_sustained_elapsed,lts__t_sector_hit_rate.pct,launch__registers_per_thread,sm__warps_active.avg.pct_of_peak_sustained_active,launch__shared_mem_per_block_dynamic
ncu --metrics $M --clock-control none -s $N -c $N -o /tmp/${TAG}_step python profiles/one_step.py 2 > /dev/null 2>&1
python profiles/summarize.py ncu /tmp/${TAG}_step.ncu-rep > gpurun_out/${TAG}_ncu_step_raw.md
echo $N > gpurun_out/${TAG}_launches_per_step.txt
