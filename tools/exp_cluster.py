"""Development: per-iteration time of the cluster SMO kernel on 10 config-2 sub-problems, for env-selected variants."""
import sys, os, subprocess
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
variants = [
    ("single", {"B200GS_SMO_CLUSTER": "0"}),
    ("co cl2 512x8", {"B200GS_SMO_CLUSTER": "2"}),
    ("co cl4 512x4", {"B200GS_SMO_CLUSTER": "4"}),
    ("co cl4 256x8", {"B200GS_SMO_CLUSTER": "4", "B200GS_SMO_NT": "256"}),
    ("co cl8 256x4", {"B200GS_SMO_CLUSTER": "8"}),
    ("co cl8 512x2", {"B200GS_SMO_CLUSTER": "8", "B200GS_SMO_NT": "512"}),
    ("co1024 cl8 1024x1", {"B200GS_SMO_CLUSTER": "8", "B200GS_SMO_NT": "1024"}),
    ("co1024 cl4 1024x2", {"B200GS_SMO_CLUSTER": "4", "B200GS_SMO_NT": "1024"}),
    ("co1024 cl2 1024x4", {"B200GS_SMO_CLUSTER": "2", "B200GS_SMO_NT": "1024"}),
]
sel = sys.argv[1:] 
for name, env in variants:
    if sel and not any(s in name for s in sel): continue
    e = dict(os.environ); e.update(env); e["B200GS_SMO_CLUSTER_N"] = "100000"
    for prof in ("0", "1"):
        e["B200GS_SMO_PROF"] = prof
        r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "exp_one.py")], env=e, capture_output=True, text=True, timeout=300)
        out = (r.stdout + r.stderr).strip().splitlines()
        keep = [l for l in out if "us/iter" in l or "prof]" in l]
        print("%-22s prof=%s | %s" % (name, prof, " || ".join(k[-230:] for k in keep) if keep else out[-3:]), flush=True)
