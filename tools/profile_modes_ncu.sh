set -e
for M in "$@"; do
  python tools/profile_modes.py $M && ncu --set full --clock-control none --import-source on -k regex:ring_step_kernel -s 3 -c 1 -f -o /tmp/$M python tools/profile_modes.py $M > /tmp/ncu_$M.log 2>&1 && ncu -i /tmp/$M.ncu-rep --page raw --csv > gpurun_out/ncu_${M}_raw.csv && ncu -i /tmp/$M.ncu-rep --page source --csv > gpurun_out/ncu_${M}_src.csv
done
ls -la gpurun_out | tail -8
