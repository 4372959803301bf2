"""C5 with the SDF solid / the Volume switched off: where the marching time goes (development)."""
import sys
sys.path.insert(0, '/root/repo')
from ptsharp_b200 import scenes
from ptsharp_b200.bindings import HostWorld, Device
dev = Device(0)
for kw in (dict(), dict(with_sdf=False), dict(with_volume=False), dict(with_sdf=False, with_volume=False)):
    hw = HostWorld(); cfg = scenes.build_c5(hw, **kw); dev.upload(hw)
    W, H, spp = cfg.width, cfg.height, 2
    dev.render_pass(hw.make_pass(W, H, spp, pass_index=0), want_mean=False)
    dev.reset_counters(); dev.render_pass(hw.make_pass(W, H, spp, pass_index=1), want_mean=False); c = dev.counters()
    print(kw, f"{c['lastPassMs']:.1f} ms, {c['segments']/c['lastPassMs']/1e6:.3f} Gseg/s, segments {c['segments']}, shadow {c['shadowRays']}")
