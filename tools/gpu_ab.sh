mkdir -p gpurun_out
for v in default bn128 s4 bn128s6; do
    if [ $v = default ]; then unset SLIDE_B200_LIB; else export SLIDE_B200_LIB=$PWD/slide_b200/libslide_b200_$v.so; fi
    python tools/profile_records.py lat 256 auto > gpurun_out/ab_lat_${v}.txt 2>&1
    echo "$v: $(head -1 gpurun_out/ab_lat_${v}.txt)"
    grep -E "SA1.att.v |SA1.att.w2|SA1.mlp.res |SA1.mlp.conv2 |SA1.att.w1k " gpurun_out/ab_lat_${v}.txt | cut -c1-100
done
