set -u
O=gpurun_out; mkdir -p $O
for p in 128 0 64 256; do DECO_ATTN_L2PROMO=$p python scripts/attn_pitch_bench.py; done > $O/attn_pitch_r2.txt 2>&1
cat $O/attn_pitch_r2.txt
for p in 128 0; do
DECO_ATTN_L2PROMO=$p ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:attention_tc -c 46 --csv --log-file $O/attn_pitch_ncu_$p.csv python scripts/attn_pitch_bench.py > /dev/null 2>&1
python - <<PY
import csv
rows=list(csv.DictReader([l for l in open("gpurun_out/attn_pitch_ncu_$p.csv") if not l.startswith("==")]))
by={}
for r in rows:
    by.setdefault(r["ID"],{})[r["Metric Name"]]=(float(r["Metric Value"].replace(",","")),r["Metric Unit"])
ids=sorted(by,key=int)
for name,sel in (("pitch 72",ids[5:20]),("pitch 80",ids[28:43])):
    rd=sum(by[i]["dram__bytes_read.sum"][0] for i in sel)/len(sel); u=by[sel[0]]["dram__bytes_read.sum"][1]
    t=sum(by[i]["gpu__time_duration.sum"][0] for i in sel)/len(sel); tu=by[sel[0]]["gpu__time_duration.sum"][1]
    print("promo $p", name, "dram read", rd, u, "time", t, tu)
PY
done
