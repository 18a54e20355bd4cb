for round in 1 2; do
for k in v1 ws; do echo -n "decoder $k: "; timeout 300 python scripts/kernel_time.py cfg2,cfg3,cfg4 20 NDZB_DECOMPRESS_KERNEL=$k 2>&1 | sed -E "s/.*(cfg[0-9]) compress avg ([0-9.]+) min [0-9.]+ ms frac ([0-9.]+) \| decompress avg ([0-9.]+) min [0-9.]+ ms frac ([0-9.]+) \| (.*)/\1 d \4 (\5) \6;/" | tr '\n' ' '; echo; done
done
