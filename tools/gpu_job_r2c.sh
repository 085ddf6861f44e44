#!/bin/bash
# r02 call C: ncu --set full of the tile kernel (variant 6) at the bench size, with source.
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_forces_fluid_tile -s 2 -c 1 -f -o gpurun_out/r02c_forces_tile_10M python tools/profile_forces.py 10000000 6 3 > gpurun_out/r02c_prof.log 2>&1
tail -2 gpurun_out/r02c_prof.log
