#!/bin/bash
# DRAM traffic and speed of one 1Mi-token forward launch as a function of the pinned L2 fraction of the weight stream.
OUT=gpurun_out; mkdir -p $OUT
for f in ${@:-1.0 0.7 0.55 0.4 0.25}; do
  echo "== RQAE_L2_HOT=$f"
  RQAE_L2_HOT=$f timeout 300 python tools/prof_forward.py --tokens 1048576 --reps 1 2>&1 | tail -1
  RQAE_L2_HOT=$f timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct --clock-control none -k regex:rq_forward -s 1 -c 1 --csv \
    --log-file $OUT/l2sweep_$f.csv python tools/prof_forward.py --tokens 1048576 --reps 1 > /dev/null 2>&1
  python tools/traffic_json.py $OUT/l2sweep_$f.csv 1048576; grep hit_rate $OUT/l2sweep_$f.csv | cut -d, -f12-
done
