cd /root/repo
O=gpurun_out; mkdir -p $O; : > $O/ab_r3.log
for round in 1 2; do
  for v in base pf3 pf3d3 wl wlst1 ldef ldef2 all; do
    echo "== $v (round $round)" >> $O/ab_r3.log
    NERFQ_LIB=profiles/_ab/$v/libnerfq.so timeout 120 python profiles/time_mlp.py >> $O/ab_r3.log 2>&1
  done
done
cat $O/ab_r3.log
