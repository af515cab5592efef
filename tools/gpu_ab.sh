mkdir -p gpurun_out
for v in default r96 r128; do
  for f in 1 0; do
    if [ $v = default ]; then unset SLIDE_B200_LIB; else export SLIDE_B200_LIB=$PWD/slide_b200/libslide_b200_$v.so; fi
    SLIDE_FUSE_SOFTMAX=$f python tools/profile_records.py lat 256 auto > gpurun_out/ab_lat_${v}_f$f.txt 2>&1
    echo "$v fuse=$f: $(head -1 gpurun_out/ab_lat_${v}_f$f.txt)"
    grep -E "SA1.att.v |SA1.att.w2|SA1.att.softmax|SA1.mlp.res |SA1.mlp.conv1 " gpurun_out/ab_lat_${v}_f$f.txt | cut -c1-100
  done
done
unset SLIDE_B200_LIB
python tools/profile_records.py pos 256 auto | head -1
SLIDE_B200_LIB=$PWD/slide_b200/libslide_b200_r96.so python tools/profile_records.py pos 256 auto | head -1
