import csv, re, sys
lines = [l for l in open(sys.argv[1] if len(sys.argv) > 1 else 'gpurun_out/launches.csv') if not l.startswith('==')]
rows = list(csv.DictReader(lines))
names = [r['Kernel Name'] for r in rows]; durs = [float(r['Metric Value'].replace(',', '')) for r in rows]
idx = [i for i, n in enumerate(names) if n.startswith('k_measure')]
# the first full step after warm-up: second k_measure
start = idx[1] - 3 if len(idx) > 1 else idx[0] - 3
end = idx[2] - 3 if len(idx) > 2 else len(names)
tot = 0
for n, d in zip(names[start:end], durs[start:end]):
    print(f"{d / 1e3:8.1f}  {re.sub(r'\(.*', '', n)}"); tot += d
print(f"{tot / 1e3:8.1f}  TOTAL ({end - start} launches)")
