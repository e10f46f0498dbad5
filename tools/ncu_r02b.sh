#!/bin/bash
# Round 2, second half: ncu --set full captures of the kernels added after tools/ncu_r02.sh (EPI_DELTA dX GEMM, LayerNorm +
# row contraction) and the launch list of the bench command with the final defaults.  Run under gpurun.
set -u
O=gpurun_out
NCU="ncu --set full --clock-control none --import-source on -f"
$NCU -k regex:gemm_cp_kernel -s 7 -c 1 -o $O/r02b_gemm_projdx_delta_N768_K768 python tools/gemm_one.py 50432 768 768 3 1 > $O/ncu_projdx_delta.log 2>&1
LN_ONCE=1 $NCU -k regex:ln_fwd_rows\|ln_bwd_rows -c 2 -o $O/r02b_ln_rows python tools/ln_rows_time.py > $O/ncu_ln_rows.log 2>&1
ATTN_TIME=0 $NCU -k regex:attn_bwd_tc -s 1 -c 1 -o $O/r02b_attn_bwd python tools/attn_one.py > $O/ncu_attn_bwd.log 2>&1
python tools/ncu_summary.py $O/r02b_gemm_projdx_delta_N768_K768.ncu-rep $O/r02b_ln_rows.ncu-rep $O/r02b_attn_bwd.ncu-rep > $O/r02b_ncu_full_summary.txt 2>&1
# launch list of the bench command: every launch, the summary takes the last two replayed steps
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/r02b_launches_all.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $O/ncu_launches_b.log 2>&1
python - <<'PY'
import csv
src = "gpurun_out/r02b_launches_all.csv"
lines = [l for l in open(src) if l.startswith('"')]
hdr, rows = lines[0], lines[1:]
# the bench replays the step graph 3 (warm-up) + 2 (timed) times: the replays are the tail of the list; one step =
# the launches between two consecutive adamw_dev_kernel launches
idx = [i for i, l in enumerate(rows) if "adamw_dev_kernel" in l]
per_step = idx[-1] - idx[-2]
tail = rows[idx[-3] + 1: idx[-1] + 1]
open("gpurun_out/r02b_launches.csv", "w").write(hdr + "".join(tail))
print("launches per step:", per_step, " kept:", len(tail), " of", len(rows))
PY
python tools/launch_summary.py $O/r02b_launches.csv 40 > $O/r02b_launches_summary.txt
rm -f $O/r02b_launches_all.csv
python tools/kernel_times.py > $O/r02b_kernel_times.log 2>&1
ls -la $O/r02b_*
