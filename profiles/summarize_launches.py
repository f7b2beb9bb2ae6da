import csv, collections, re, sys
lines=[l for l in open(sys.argv[1]) if not l.startswith('==')]
tot={}; cnt=collections.Counter()
for row in csv.DictReader(lines):
    name=re.sub(r'\(.*','',row['Kernel Name']); name=re.sub(r'^void ','',name)
    v=float(row['Metric Value'].replace(',','')); u=row['Metric Unit']
    v = v/1e3 if u=='ns' else v*1e3 if u=='ms' else v
    tot[name]=tot.get(name,0)+v; cnt[name]+=1
T=sum(tot.values())
print(f"total kernel time {T/1e3:.2f} ms over {sum(cnt.values())} launches")
print("| kernel | launches | total us | share | avg us |\n|---|---|---|---|---|")
for k,v in sorted(tot.items(), key=lambda kv:-kv[1])[:int(sys.argv[2]) if len(sys.argv)>2 else 18]:
    print(f"| `{k[:80]}` | {cnt[k]} | {v:.0f} | {100*v/T:.1f}% | {v/cnt[k]:.1f} |")
