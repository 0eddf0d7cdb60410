"""GPU: time of one tcgen05 gradient launch (config 3 data) as a function of the number of listed
chains and the row cut.  With no argument: sweep (one subprocess per forced chunk count, the env
knob TNUTS_TC_CHUNKS is read once per process); `--one` prints the row for this process."""
import json, os, subprocess, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def one():
    import turing_jl_b200 as T
    model = T.models.config3_logistic()
    out = {}
    for C in (4096, 3968, 2048, 1280, 1152, 896, 640, 512, 384, 256, 128, 16):
        e = T.Engine(model, T.NUTS(0.8), C, 0, seed=1)
        e.set_option("logistic_tc", 1)
        e.init()
        e.bench_gradient(3)
        out[C] = round(e.bench_gradient(20) * 1e6, 1)
        e.close()
    print(json.dumps({"chunks": os.environ.get("TNUTS_TC_CHUNKS", "model"), "us_per_gradient_launch": out}))


if __name__ == "__main__":
    if "--one" in sys.argv:
        one()
    else:
        for c in ("", "9", "18", "37", "74", "148"):
            env = dict(os.environ)
            if c:
                env["TNUTS_TC_CHUNKS"] = c
            subprocess.run([sys.executable, os.path.abspath(__file__), "--one"], env=env, check=False)
