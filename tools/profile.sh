#!/bin/bash
# ncu evidence for profiles/ (run on the GPU box through gpurun): tools/profile.sh <tag>
# Launch lists (gpu__time_duration only, every nis kernel of one whole step / one whole query) and one --set full capture of two
# tracking batches and three scan batches.  The .ncu-rep files stay in /tmp on the box (3 MB per kernel with --import-source on);
# only the raw-page and source-page CSVs come back in gpurun_out/.
tag=${1:-r02}
what=${2:-all}          # all | track | scan
K='regex:^(col_|colcol|row|polar_|rzc_|pose_|select_|scan_|transpose)'
B="python bench.py --db 0 --cpu-frames 0 --no-extras --no-cfg4 --steps 1 --warmup 1"
mkdir -p gpurun_out
[ "$what" = scan ] || ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -s 396 -c 396 --csv --log-file gpurun_out/launches_${tag}.csv $B > gpurun_out/ncu_${tag}_l.log 2>&1
[ "$what" = track ] || ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k "$K" -s 480 -c 1400 --csv --log-file gpurun_out/launches_${tag}_scan.csv python tools/scan_probe.py 2048 2 > gpurun_out/ncu_${tag}_sl.log 2>&1
[ "$what" = scan ] || ncu --set full --clock-control none --import-source on -k "$K" -s 418 -c 44 -o /tmp/prof_track $B > gpurun_out/ncu_${tag}_t.log 2>&1
[ "$what" = scan ] || ncu -i /tmp/prof_track.ncu-rep --page raw --csv > gpurun_out/raw_${tag}_track.csv
[ "$what" = scan ] || ncu -i /tmp/prof_track.ncu-rep --page source --csv --print-source sass | gzip > gpurun_out/source_${tag}_track.csv.gz
[ "$what" = track ] || ncu --set full --clock-control none --import-source on -k "$K" -s 1300 -c 45 -o /tmp/prof_scan python tools/scan_probe.py 2048 2 > gpurun_out/ncu_${tag}_s.log 2>&1
[ "$what" = track ] || ncu -i /tmp/prof_scan.ncu-rep --page raw --csv > gpurun_out/raw_${tag}_scan.csv
[ "$what" = track ] || ncu -i /tmp/prof_scan.ncu-rep --page source --csv --print-source sass | gzip > gpurun_out/source_${tag}_scan.csv.gz
du -sh gpurun_out
