set -u
mkdir -p gpurun_out
# 1) launch list of the default bench (N=1, cfg5) — only after the same command exited 0 without ncu
CMD="python bench.py --steps 5 --warmup 3 --no-cpu-baseline --quick"
timeout 300 $CMD > gpurun_out/plain.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launch.log 2>&1
echo "ncu launches exit=$?"
# 2) full capture of the scan kernel at 10 M rows
timeout 300 $CMD > gpurun_out/plain2.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:scan_topk -s 3 -c 1 -o gpurun_out/prof_scan -f $CMD > gpurun_out/ncu_full.log 2>&1
echo "ncu full exit=$?"
# 3) full capture at one of eight shards (1.25 M rows)
CMD8="python bench.py --bank-rows 1250000 --steps 5 --warmup 3 --no-cpu-baseline --quick"
timeout 300 $CMD8 > gpurun_out/plain_s8.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:scan_topk -s 3 -c 1 -o gpurun_out/prof_scan_shard8 -f $CMD8 > gpurun_out/ncu_s8.log 2>&1
echo "ncu shard8 exit=$?"
# 4) k+s=16 and 32 at shard size
CMDK="python tools/probe_one.py 128 1250000 512 32"
timeout 300 $CMDK > gpurun_out/plain_k32.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:scan_topk -s 2 -c 1 -o gpurun_out/prof_k32 -f $CMDK > gpurun_out/ncu_k32.log 2>&1
echo "ncu k32 exit=$?"
python tools/show_bench.py gpurun_out/plain.log; python tools/show_bench.py gpurun_out/plain_s8.log; tail -1 gpurun_out/plain_k32.log
