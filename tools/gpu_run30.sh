mkdir -p gpurun_out /tmp/ncu
timeout 900 ncu --set full --clock-control none --import-source on -k regex:bsq_k3_walk -c 1 -o /tmp/ncu/k3walk python tools/kbench.py c5_int 1 1 > gpurun_out/r02D_full_walk.log 2>&1
python profiles/summarize_ncu.py /tmp/ncu/k3walk.ncu-rep > gpurun_out/r02D_k3walk_ncu.txt 2>&1
ncu -i /tmp/ncu/k3walk.ncu-rep --page source --csv 2>/dev/null | gzip > gpurun_out/r02D_k3walk_source.csv.gz
head -30 gpurun_out/r02D_k3walk_ncu.txt
