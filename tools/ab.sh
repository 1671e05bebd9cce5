# A/B of two builds of the library on full-size single layers (one process per case): tools/ab.sh case...
for lib in libdispnet_b200_base.so libdispnet_b200.so; do
  echo "== $lib"
  for c in "$@"; do
    DISPNET_B200_LIB=$PWD/supervised_dispnet_b200/$lib timeout 120 python tools/prof_conv.py $c 2>&1 | grep " tc \| cc "
  done
done
