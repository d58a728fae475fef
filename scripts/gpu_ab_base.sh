bash scripts/gpu_suite.sh r02x
for mode in 0 1; do
  for wl in "c3 148" "c1 444"; do
    MODLE_B200_LIB=$PWD/modle_b200/libmodle_b200_base.so MODLE_B200_RNG_MODE=$mode timeout 200 python scripts/gpu_phases.py $wl
  done > gpurun_out/r02x_base_phases_mode$mode.txt 2>&1
  grep libmodle gpurun_out/r02x_base_phases_mode$mode.txt
done
