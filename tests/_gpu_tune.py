"""Scheduler tuning sweep (env overrides are read once per process)."""
import os, subprocess, sys
name = sys.argv[1] if len(sys.argv) > 1 else "headline"
n = sys.argv[2] if len(sys.argv) > 2 else "400000"
configs = [dict(PUMAS_B200_WF_SLOTS=s, PUMAS_B200_WF_EHS=e, PUMAS_B200_WF_DEL=d)
           for (s, e, d) in [(64, 16, 8), (64, 32, 8), (64, 1, 8), (96, 32, 8), (48, 16, 8)]]
for c in configs:
    env = dict(os.environ); env.update({k: str(v) for k, v in c.items()}); env["PUMAS_B200_WF_STATS"] = "1"
    r = subprocess.run([sys.executable, "tests/_gpu_prof.py", name, n, "2"], env=env, capture_output=True, text=True)
    last = [l for l in r.stdout.splitlines() if "tracks/s" in l][-1]
    st = [l for l in r.stderr.splitlines() if "[wavefront]" in l][-1]
    print(c["PUMAS_B200_WF_SLOTS"], c["PUMAS_B200_WF_EHS"], c["PUMAS_B200_WF_DEL"], "|", last.split("->")[1].strip(), "|", st.replace("[wavefront] ", ""), flush=True)
