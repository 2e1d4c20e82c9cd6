"""A/B of library builds: for each .so given, run workloads in a subprocess."""
import os, subprocess, sys
libs = sys.argv[1].split(",")
jobs = sys.argv[2:] or ["headline:1000000"]
for lib in libs:
    for job in jobs:
        name, n = job.split(":")
        env = dict(os.environ); env["PUMAS_B200_LIBRARY"] = os.path.abspath(lib); env["PUMAS_B200_WF_STATS"] = "1"
        for extra in (os.environ.get("AB_ENV", "").split(";")):
            if "=" in extra:
                k, v = extra.split("="); env[k] = v
        r = subprocess.run([sys.executable, "tests/_gpu_prof.py", name, n, "2"], env=env, capture_output=True, text=True)
        out = [l for l in r.stdout.splitlines() if "tracks/s" in l]
        st = [l for l in r.stderr.splitlines() if "[wavefront]" in l]
        print(os.path.basename(lib), job, "|", out[-1].split("kernel")[1].strip() if out else r.stderr[-300:], "|", st[-1][12:] if st else "", flush=True)
