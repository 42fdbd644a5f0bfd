cp sigkernel_b200/libsigkernel_b200.so /tmp/lib_new.so
for i in 1 2 3; do
  cp /tmp/lib_new.so sigkernel_b200/libsigkernel_b200.so
  echo -n "new: "; SKB_WPSM=0 timeout 300 python tools/time_fwd.py cfg3 2>&1 | grep rbf
  cp tools/lib_before.so sigkernel_b200/libsigkernel_b200.so
  echo -n "old: "; SKB_WPSM=0 timeout 300 python tools/time_fwd.py cfg3 2>&1 | grep rbf
done
cp /tmp/lib_new.so sigkernel_b200/libsigkernel_b200.so
