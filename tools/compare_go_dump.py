"""Compare a dump written by go/harness/cubedrop_headless.go (the unmodified Go reference) with
the committed golden fixture / the oracle: contact counts, pair-sequence hashes per frame and the
raw bits of the final state.  This is how the oracle<->Go gap gets closed on a machine with Go."""
import os, re, struct, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
gold = np.load(os.path.join(ROOT, "tests", "golden", "cubedrop_f64.npz"))
steps, bodies = [], {}
for line in open(sys.argv[1]):
    m = re.match(r"step (\d+) contacts (\d+) pairhash ([0-9a-f]+)", line)
    if m:
        steps.append((int(m.group(2)), int(m.group(3), 16)))
    m = re.match(r"body (\d+) pos (.*) awake", line)
    if m:
        bodies[int(m.group(1))] = line
n = min(len(steps), gold["counts"].shape[0])
bad = [s for s in range(n) if steps[s][0] != int(gold["counts"][s, 0]) or steps[s][1] != int(gold["pair_hash"][s, 0])]
print(f"{n} frames compared; first contact-set mismatch: {bad[0] if bad else None}")
if len(steps) == gold["counts"].shape[0]:
    def bits(x): return struct.unpack("<Q", struct.pack("<d", float(x)))[0]
    ok = True
    for i, line in bodies.items():
        got = [int(t, 16) for t in re.findall(r"\b[0-9a-f]{1,16}\b", line.split("pos", 1)[1].split("awake")[0])]
        want = [bits(v) for v in list(gold["position"][i]) + list(gold["orientation"][i]) + list(gold["velocity"][i]) + list(gold["rotation"][i])]
        ok &= got == want
    print("final state bit-identical to the oracle golden:", ok)
