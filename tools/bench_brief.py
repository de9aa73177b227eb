"""Development aid: one-line summary of a bench.py JSON line."""
import json, sys
for f in sys.argv[1:]:
    d = json.loads(open(f).read().strip().splitlines()[-1])
    k = d['kernel_ms']; ks = list(k.values())
    o = d.get('other_parity_mode')
    print(f"{f}: {d['config']['parity']} {d['ms_per_step']:.3f} ms/step (step {ks[0]:.3f} obs {ks[1]:.3f} side {ks[2]:.3f} rest {ks[3]:.3f}) "
          f"obs frac {d['roofline']['frac']:.3f} whole {d['roofline']['whole_step']['frac']:.3f}"
          + (f" | {o['parity']} {o['ms_per_step']:.3f} (step {list(o['kernel_ms'].values())[0]:.3f} obs {list(o['kernel_ms'].values())[1]:.3f}) whole {o['roofline']['whole_step']['frac']:.3f}" if o else ''))
