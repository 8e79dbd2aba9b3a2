"""Time the E-step passes of every library variant under variants/ (scripts/build_variants.sh) on one GPU.
One subprocess per variant (PMM_LIB is read when the library is first loaded); the workload is generated once and
cached under /tmp.  Prints one line per (variant, workload): inside / outside / E-step ms and the outside pass's
fraction of the measured HBM peak on the bytes the schedule must move.
    python scripts/exp_variants.py [--only name,name] [--workloads config4,shard250,config3]"""
import argparse, glob, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def worker(workloads):
    import numpy as np
    from laml_b200.capi import Engine
    from laml_b200.synth import CONFIGS, make_workload
    peak = 6451.8
    try:
        peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        pass
    for wl in workloads:
        cache = "/tmp/pmm_wl_%s.npz" % wl
        spec = dict(CONFIGS[{"config3": "config3_10k_x_500", "config2": "config2_1k_x_30", "config5": "config5_5k_x_300"}.get(wl, "config4_100k_x_2k")])
        if wl.startswith("shard"):
            spec["n_sites"] = int(wl[5:])
        if os.path.exists(cache):
            z = np.load(cache)
            from laml_b200.flatten import forest_from_parent
            f = forest_from_parent(z["parent"], z["left"], z["right"])
            d, qptr, qval, codes = z["d"], z["qptr"], z["qval"], z["codes"]
        else:
            w = make_workload(seed=1, device="cuda", **spec)
            f, d, qptr, qval, codes = w.forest, w.d, w.qptr, w.qval, w.codes
            np.savez(cache, parent=f.parent, left=f.left, right=f.right, d=d, qptr=qptr, qval=qval, codes=codes)
        N, K = f.n_leaves, codes.shape[1]
        eng = Engine(N, qptr, qval, codes)
        eng.set_tree(f.parent, f.left, f.right)
        eng.set_option("slices", 1)
        eng.upload_params(d, spec["nu"], spec["phi"])
        for _ in range(3):
            eng.estep_dev()
        eng.sync()
        ts = []
        for _ in range(8):
            eng.estep_dev()
            ts.append(eng.last_timing())
        ins, out, tot = (float(np.median([t[k] for t in ts])) for k in ("inside_ms", "outside_ms", "total_ms"))
        eng.set_option("slices", int(os.environ.get("EXP_SLICES", "2")))
        for _ in range(2):
            eng.estep_dev()
        t2 = []
        for _ in range(8):
            eng.estep_dev()
            t2.append(eng.last_timing()["total_ms"])
        info = eng.info()
        st = eng.estep(d, spec["nu"], spec["phi"])
        print(json.dumps({"workload": wl, "inside_ms": ins, "outside_ms": out, "estep_ms": tot, "estep_2slices_ms": float(np.median(t2)),
                          "out_frac": info["outside_bytes_per_site"] * K / (out * 1e-3) / 1e9 / peak,
                          "in_frac": info["inside_bytes_per_site"] * K / (ins * 1e-3) / 1e9 / peak,
                          "clades": info["clades"], "llh": st["llh"], "S_sum": float(st["S"].sum())}), flush=True)
        eng.close()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default="")
    ap.add_argument("--workloads", default="config4,shard250,config3")
    ap.add_argument("--worker", action="store_true")
    a = ap.parse_args()
    wls = a.workloads.split(",")
    if a.worker:
        return worker(wls)
    libs = sorted(glob.glob(os.path.join(ROOT, "variants", "libpmm_*.so")))
    if a.only:
        keep = set(a.only.split(","))
        libs = [l for l in libs if os.path.basename(l)[7:-3] in keep]
    ref = {}
    for lib in libs:
        name = os.path.basename(lib)[7:-3]
        res = subprocess.run([sys.executable, os.path.abspath(__file__), "--worker", "--workloads", a.workloads],
                             env=dict(os.environ, PMM_LIB=lib), capture_output=True, text=True, timeout=900)
        if res.returncode != 0:
            print("%-14s FAILED: %s" % (name, res.stderr.strip().splitlines()[-1:] or res.stderr[-300:]), flush=True)
            continue
        for line in res.stdout.splitlines():
            try:
                r = json.loads(line)
            except ValueError:
                continue
            key = r["workload"]
            same = ""
            if key in ref:
                same = " llh %s S %.1e" % ("same" if r["llh"] == ref[key][0] else "DIFF %.2e" % abs(r["llh"] / ref[key][0] - 1),
                                           abs(r["S_sum"] / ref[key][1] - 1))
            else:
                ref[key] = (r["llh"], r["S_sum"])
            print("%-14s %-9s inside %.3f (%.3f) outside %.3f (%.3f) estep %.3f  2-slices %.3f ms%s" % (
                name, key, r["inside_ms"], r["in_frac"], r["outside_ms"], r["out_frac"], r["estep_ms"], r["estep_2slices_ms"], same), flush=True)


if __name__ == "__main__":
    main()
