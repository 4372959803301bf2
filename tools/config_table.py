"""Throughput of every BASELINE config at full resolution and scene size, reduced spp (one GPU).  Development tool;
writes a markdown table to stdout.  Usage: python tools/config_table.py [c1 c2 c3 c4 c5] [--spp N]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ptsharp_b200 import scenes
from ptsharp_b200.bindings import HostWorld, Device

args = [a for a in sys.argv[1:] if not a.startswith('--')]
spp_override = None
if '--spp' in sys.argv:
    spp_override = int(sys.argv[sys.argv.index('--spp') + 1]); args = [a for a in args if a != str(spp_override)]
names = args or ['c1', 'c2', 'c3', 'c4', 'c5']
DEFAULT_SPP = {'c1': 16, 'c2': 32, 'c3': 8, 'c4': 2, 'c5': 2}
print("| config | resolution | spp (this run) | triangles | build s | pass ms | Msamples/s | Gpaths·bounce/s | shadow Grays/s | trace/shade/shadow ms | of which k_mesh / k_march<SDF> / k_march<VOLUME> ms |")
print("|---|---|---|---|---|---|---|---|---|---|---|")
for n in names:
    hw = HostWorld()
    t0 = time.time(); cfg = scenes.BUILDERS[n](hw); flat = hw.flatten(); build = time.time() - t0
    dev = Device(0); dev.upload_flat(flat)
    spp = spp_override or DEFAULT_SPP[n]
    dev.render_pass(hw.make_pass(cfg.width, cfg.height, spp, pass_index=0), want_mean=False)  # warm-up at the measured spp: queues grow on demand
    dev.reset_counters()
    dev.render_pass(hw.make_pass(cfg.width, cfg.height, spp, pass_index=1), want_mean=False)
    c = dev.counters(); ms = c['lastPassMs']
    dev.set_profiling(True)
    dev.render_pass(hw.make_pass(cfg.width, cfg.height, max(1, spp // 2), pass_index=2), want_mean=False)
    p = dev.counters()
    print(f"| {cfg.name} | {cfg.width}x{cfg.height} | {spp} of {cfg.spp} | {cfg.triangles} | {build:.1f} | {ms:.1f} | {c['cameraSamples']/ms/1e3:.1f} | "
          f"{c['segments']/ms/1e6:.3f} | {c['shadowRays']/ms/1e6:.3f} | {p['traceMs']:.0f}/{p['shadeMs']:.0f}/{p['shadowMs']:.0f} | {p['meshMs']:.0f}/{p['sdfMs']:.0f}/{p['volumeMs']:.0f} |", flush=True)
    dev.close()
