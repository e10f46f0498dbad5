#!/bin/bash
# Round 2: one ncu --set full capture per kernel at the bench shapes (ViT-B/16, batch 256: M = 50432, rank 16), plus the
# launch list of the bench command.  Run under gpurun; reports land in gpurun_out/, summaries via tools/ncu_summary.py.
set -u
O=gpurun_out
NCU="ncu --set full --clock-control none --import-source on -f"
g() { # name M N K epi
  $NCU -k regex:gemm_cp_kernel -s 7 -c 1 -o $O/r02_gemm_$1 python tools/gemm_one.py $2 $3 $4 $5 1 > $O/ncu_$1.log 2>&1
}
g fc2_N768_K3072 50432 768 3072 0
g fc1_gelu_N3072_K768 50432 3072 768 1
g fc2dx_dgelu_N3072_K768 50432 3072 768 2
g qkv_N2304_K768 50432 2304 768 0
g qkvdx_N768_K2304 50432 768 2304 0
g proj_N768_K768 50432 768 768 0
$NCU -k regex:attn_ -s 3 -c 3 -o $O/r02_attn python tools/attn_one.py > $O/ncu_attn.log 2>&1
SKINNY_ONCE=1 $NCU -k regex:rows_kernel\|cols_kernel -c 13 -o $O/r02_skinny python tools/skinny_time.py > $O/ncu_skinny.log 2>&1
$NCU -k regex:ln_ -s 4 -c 2 -o $O/r02_ln python tools/ln_one.py > $O/ncu_ln.log 2>&1
$NCU -k regex:merge_kernel\|adamw_dev_kernel\|patchify16_kernel\|assemble_kernel\|factor_operands_kernel -s 5 -c 5 -o $O/r02_misc python tools/misc_one.py > $O/ncu_misc.log 2>&1
CARA_SIDE_TILES=1 $NCU -k regex:gemm_cp_kernel -s 2 -c 1 -o $O/r02_gemm_side_qkv python tools/gemm_side_time.py > $O/ncu_side.log 2>&1
python tools/ncu_summary.py $O/r02_gemm_fc1_gelu_N3072_K768.ncu-rep $O/r02_gemm_fc2_N768_K3072.ncu-rep $O/r02_gemm_fc2dx_dgelu_N3072_K768.ncu-rep \
  $O/r02_gemm_proj_N768_K768.ncu-rep $O/r02_gemm_qkv_N2304_K768.ncu-rep $O/r02_gemm_qkvdx_N768_K2304.ncu-rep $O/r02_attn.ncu-rep \
  $O/r02_skinny.ncu-rep $O/r02_ln.ncu-rep $O/r02_misc.ncu-rep > $O/r02_ncu_full_summary.txt 2>&1
# launch list of the bench command (cold-cache, serialised: compare SHARES); skip model build + eager warm-up + 3 replays
ncu --metrics gpu__time_duration.sum --clock-control none -s 2820 -c 1000 --csv --log-file $O/r02_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $O/ncu_launches.log 2>&1
rm -f $O/r02_gemm_side_qkv.ncu-rep
ls -la $O/r02_* | head -30
