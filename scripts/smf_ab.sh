for rep in 1 2; do
for lib in "" build/lib_lh.so; do
for n in 8192 65536; do
echo "== rep=$rep lib=${lib:-default} n=$n"; LITE_B200_LIB=$lib python scripts/smf_probe.py $n 1020 1500 300 2>&1 | tail -3 | sed 's/.*| //'
done; done; done
