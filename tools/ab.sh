# A/B of two builds of the library on full-size single layers (one process per case): tools/ab.sh case...
# libdispnet_b200_base.so = the build to compare against, e.g.
#   git archive <commit> supervised_dispnet_b200/csrc include | tar -x -C /tmp/base && make -C /tmp/base/supervised_dispnet_b200/csrc \
#     && cp /tmp/base/supervised_dispnet_b200/libdispnet_b200.so supervised_dispnet_b200/libdispnet_b200_base.so
for lib in libdispnet_b200_base.so libdispnet_b200.so; do
  echo "== $lib"
  for c in "$@"; do
    DISPNET_B200_LIB=$PWD/supervised_dispnet_b200/$lib timeout 120 python tools/prof_conv.py $c 2>&1 | grep " tc \| cc "
  done
done
