#!/bin/bash
# Round-2 ncu captures (run on the GPU box through gpurun): launch lists of the default bench command and of one training
# step, `ncu --set full` of the convolution at every level, of the other kernels of levels 0-1 and of the weight
# gradient.  The full reports stay on the box (they exceed what gpurun copies back); the per-launch summaries
# (profiles/summarize_full.py) come back in gpurun_out/ and are committed under profiles/.
mkdir -p gpurun_out
cap() {   # name, ncu selection args..., --, command
  name=$1; shift
  sel=()
  while [ "$1" != "--" ]; do sel+=("$1"); shift; done
  shift
  ncu --profile-from-start off --set full --clock-control none --import-source on "${sel[@]}" -o /tmp/$name -f "$@" > gpurun_out/$name.log 2>&1
  ncu -i /tmp/$name.ncu-rep --page raw --csv 2>/dev/null | python profiles/summarize_full.py > gpurun_out/$name.json
}
ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 600 --csv --log-file gpurun_out/r2_launches_default_cmd.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extras > gpurun_out/r2_ncu_default.log 2>&1
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_train_launches.csv \
    python tools/ncu_train_step.py > gpurun_out/r2_ncu_train.log 2>&1
cap r2_conv_tc_levels -k regex:k_conv_tc -- python tools/ncu_sequence.py --no-stem
cap r2_level01_other -k 'regex:k_points|k_assign|k_vertices|k_scatter|k_normalize|k_splat_gather|k_clear|k_zero' -c 16 -- python tools/ncu_sequence.py
cap r2_wgrad_tc -k regex:k_wgrad_tc -- python tools/ncu_train_step.py
cp /tmp/r2_conv_tc_levels.ncu-rep gpurun_out/ 2>/dev/null
ls -la gpurun_out
