"""Copy the bench line and the ncu launch list of the last gpurun call into profiles/ and rebuild the launch table.
Usage: python tools/refresh_profiles.py   (reads gpurun_out/bench_final.json and gpurun_out/launches_r01q.csv)"""
import collections, csv, json, os, re, shutil
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
os.chdir(ROOT)
d = json.loads(open('gpurun_out/bench_final.json').read().strip().split('\n')[-1])
shutil.copy('gpurun_out/bench_final.json', 'profiles/r01_bench_c3_final.json')
shutil.copy('gpurun_out/launches_r01q.csv', 'profiles/r01_launches_final.csv')
rows = list(csv.reader(open('profiles/r01_launches_final.csv')))
hi = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
h = rows[hi]; ix = {n: i for i, n in enumerate(h)}
agg = collections.defaultdict(lambda: [0, 0.0])
for r in rows[hi + 1:]:
    if len(r) < len(h) or r[ix['Metric Name']] != 'gpu__time_duration.sum':
        continue
    name = re.sub(r'\(.*', '', r[ix['Kernel Name']]).replace('void ', '')
    if name.startswith('at::'):
        name = 'torch fill (d_sum.zero_)'
    agg[name][0] += 1; agg[name][1] += float(r[ix['Metric Value']].replace(',', '')) / 1e6
tot = sum(v[1] for v in agg.values()); n = sum(v[0] for v in agg.values())
tbl = "| kernel | launches | total ms | share |\n|---|---|---|---|\n"
for k, v in sorted(agg.items(), key=lambda x: -x[1][1]):
    tbl += f"| `{k}` | {v[0]} | {v[1]:.2f} | {v[1] / tot:.3f} |\n"
tbl += f"| all | {n} | {tot:.2f} | 1.000 |\n"
p = 'profiles/r01_launches_final_summary.md'
s = open(p).read()
i = s.index('| kernel | launches | total ms | share |'); j = s.index('(Last session of the round:')
s = s[:i] + tbl + '\n' + s[j:]
s = re.sub(r"`k_mesh` share of the step [\d.]+, average launch [\d.]+ ms \(512 spp, 64-spp batches\), [\d.]+ Msamples/s and [\d.]+ Gpaths·bounce/s device-timed, [\d.]+ Msamples/s end to end\.",
           f"`k_mesh` share of the step {d['roofline']['share_of_step']:.3f}, average launch {d['roofline']['avg_launch_ms']:.2f} ms (512 spp, 64-spp batches), "
           f"{d['value']:.1f} Msamples/s and {d['gpaths_bounce_per_s']:.3f} Gpaths·bounce/s device-timed, {d['e2e']['value']:.1f} Msamples/s end to end.", s)
open(p, 'w').write(s)
print(tbl)
print(f"{d['value']:.1f} Msamples/s, {d['gpaths_bounce_per_s']:.3f} Gpaths-bounce/s, e2e {d['e2e']['value']:.1f}, k_mesh share {d['roofline']['share_of_step']:.3f}, "
      f"roofline frac {d['roofline']['frac']:.4f}, cpu {d['cpu_baseline']['value']:.3f}, clocks {d['clocks']}")
