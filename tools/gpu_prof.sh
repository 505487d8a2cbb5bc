mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:contact_step_kernel -s 3 -c 1 -o gpurun_out/${1}_prof -f python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/${1}_ncu.log 2>&1
tail -3 gpurun_out/${1}_ncu.log
