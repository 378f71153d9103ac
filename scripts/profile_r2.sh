#!/bin/bash
# Round-2 ncu captures (GPU box, one GPU).  Every profiled command first runs plain and must exit 0.
# Numbers printed by runs under ncu are never bench values.
set -u
OUT=gpurun_out
BENCH="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-extras --no-standin --no-scaling-128v --profile-run"
K='regex:k_gram|k_resolve|k_sweep_units|k_face_zbuf|k_render_bwd|k_tex4|k_prepare|k_mse|k_maxpool|k_transform|k_composite|k_mesh_reg'

$BENCH > $OUT/r2_prof_plain.log 2>&1 || { echo "plain bench failed"; exit 1; }
# 1. every launch of one warm-up + one timed step with its device time
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $OUT/r2_launches_ncu.csv $BENCH \
    > $OUT/r2_prof_launches.log 2>&1
# 2. --set full over the libst3d launches of the warm-up step and of the timed step (~50 each)
ncu --set full --clock-control none -k "$K" -c 110 -o $OUT/r2_full $BENCH > $OUT/r2_prof_full.log 2>&1
ncu -i $OUT/r2_full.ncu-rep --page raw --csv > $OUT/r2_ncu_full_raw.csv 2> /dev/null
python scripts/ncu_select.py $OUT/r2_ncu_full_raw.csv $OUT/r2_ncu_full.csv
# 3. dense mesh: 1.5 M faces, 8 views x 1024^2, forward + texture backward
DENSE="env SUBDIV=4 SIZE=1024 REPS=2 NEED_VERTS=0 python scripts/render_only.py"
$DENSE > $OUT/r2_prof_dense_plain.log 2>&1 || { echo "plain dense render failed"; exit 1; }
ncu --set full --clock-control none -k "$K" -s 5 -c 5 -o $OUT/r2_full_dense $DENSE > $OUT/r2_prof_dense.log 2>&1
ncu -i $OUT/r2_full_dense.ncu-rep --page raw --csv > $OUT/r2_ncu_full_dense_raw.csv 2> /dev/null
python scripts/ncu_select.py $OUT/r2_ncu_full_dense_raw.csv $OUT/r2_ncu_full_dense_1p5M_faces.csv
# 4. the C3 regularisers (bob, forward + backward): two launches
REG="env MESHES=bob python scripts/mesh_reg_probe.py"
$REG > $OUT/r2_prof_reg_plain.log 2>&1 || { echo "plain regulariser probe failed"; exit 1; }
ncu --set full --clock-control none -k regex:k_mesh_reg -s 6 -c 2 -o $OUT/r2_full_reg $REG > $OUT/r2_prof_reg.log 2>&1
ncu -i $OUT/r2_full_reg.ncu-rep --page raw --csv > $OUT/r2_ncu_full_reg_raw.csv 2> /dev/null
python scripts/ncu_select.py $OUT/r2_ncu_full_reg_raw.csv $OUT/r2_ncu_full_mesh_regularisers.csv
rm -f $OUT/r2_full.ncu-rep $OUT/r2_full_dense.ncu-rep $OUT/r2_full_reg.ncu-rep $OUT/*_raw.csv
ls -la $OUT | tail -12
