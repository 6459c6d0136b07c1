#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export LATTE_B200_NO_BUILD=1
timeout 400 ncu --set full --clock-control none -k regex:"attn_bwd|ln_modulate_bwd|gate_bwd|gelu_bwd|colsum|gate_residual_ln" -s 8 -c 8 -o gpurun_out/gg_prof_train_passes -f python tools/gpu_train_ncu_probe.py > gpurun_out/gg_ncu.log 2>&1
ncu -i gpurun_out/gg_prof_train_passes.ncu-rep --page raw --csv 2>/dev/null | python -c "
import csv,sys
rows=list(csv.reader(sys.stdin))
hdr=rows[0]
want=['Kernel Name','gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','dram__throughput.avg.pct_of_peak_sustained_elapsed','sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active','sm__warps_active.avg.pct_of_peak_sustained_active','launch__registers_per_thread','lts__t_sector_hit_rate.pct']
idx=[hdr.index(w) for w in want if w in hdr]
print(' | '.join(hdr[i] for i in idx))
print(' | '.join(rows[1][i] for i in idx))
for r in rows[2:]:
    print(' | '.join(r[i][:44] for i in idx))
" > gpurun_out/gg_train_passes_ncu.txt
cat gpurun_out/gg_train_passes_ncu.txt; tail -3 gpurun_out/gg_ncu.log
timeout 300 python -m pytest tests/test_gpu_train.py -q -k "attention_backward or reference_gradients" 2>&1 | tail -2
