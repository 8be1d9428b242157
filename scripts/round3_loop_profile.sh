O=gpurun_out/r3
mkdir -p $O
timeout 300 env SWEEP_SET=r3g SWEEP_ONLY="early 4 la4 cheap" SWEEP_STEPS=30 python scripts/sweep_c3.py 2>&1 | grep -v Warn | tee $O/sweep_r3h.log
MKB_PROFILE_OPTS="dict(tile_loop=True)" timeout 300 ncu --set full --clock-control none --import-source on -k regex:mkb_cell_step -s 4 -c 1 -f -o $O/prof_c3_loop python scripts/profile_target.py c3 6 > $O/ncu_c3_loop.log 2>&1
tail -1 $O/ncu_c3_loop.log
