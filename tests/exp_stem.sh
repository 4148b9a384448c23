for f in "" "stem_dbg=4" "stem_dbg=2" "stem_dbg=6" "stem_dbg=1" "stem=2"; do
  echo "== VSB_FLAGS=$f"; VSB_FLAGS=$f timeout 300 python tests/layer_profile.py 1024 64 2>&1 | grep -E "encoder.conv1|total conv"
done
for b in 32 64 128; do
  echo "== batch $b"; timeout 300 python tests/layer_profile.py 1024 128 $b 2>&1 | grep -E "total conv|slicer"
done
