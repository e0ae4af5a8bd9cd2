"""Generate tests/golden/*.npz from the LIVE reference (/root/reference, build container only).

TEST INFRASTRUCTURE.  Run:  python oracle/make_golden.py
Every array written here is an output of the unmodified reference code (imported through
oracle/ref_harness.py), plus the inputs/draws needed to reproduce it without the reference.
The fixtures travel to the GPU box; /root/reference does not.
"""
import hashlib
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_harness as rh      # noqa: E402
import oracle_np as O         # noqa: E402  (only for draw capture helpers; outputs come from the reference)

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")


def _jsonable(p):
    d = {}
    for k, v in p.items():
        if isinstance(v, np.ndarray):
            d[k] = {"nd": v.tolist(), "dtype": str(v.dtype)}
        elif isinstance(v, np.floating):
            d[k] = {"f64": float(v).hex()}
        elif isinstance(v, float):
            d[k] = {"pyf": v.hex()}
        else:
            d[k] = {"pyi": int(v)}
    return d


def main():
    os.makedirs(OUT, exist_ok=True)
    R = rh.load()
    P, ISP = R.process, R.isp_ops
    meta = {"versions": rh.versions()}

    # ---- P1/P2: exhaustive sensor-code tables
    pack = {}
    for cam, wp, bl, n in (("sony", 16383, 512, 16384), ("imx686", 1023, 64, 1024)):
        codes = np.arange(n, dtype=np.uint16)
        raw = np.zeros((2, 2 * n), np.uint16)
        raw[:, 0::2] = codes
        raw[:, 1::2] = codes
        for clip in (False, True):
            t = ISP.raw2bayer(raw, wp=wp, bl=bl, norm=True, clip=clip)
            pack[f"{cam}_clip{int(clip)}"] = t[0, 0].copy()
            meta[f"pack_sha1_{cam}_clip{int(clip)}"] = hashlib.sha1(t[0, 0].tobytes()).hexdigest()[:16]
        pack[f"{cam}_roundtrip"] = ISP.bayer2raw(ISP.raw2bayer(raw, wp=wp, bl=bl, norm=True, clip=True), wp=wp, bl=bl)[0, 0::2]
    rng = np.random.RandomState(1)
    raw = rng.randint(0, 16384, size=(12, 16)).astype(np.uint16)
    pack["rand_raw"] = raw
    pack["rand_packed"] = ISP.raw2bayer(raw, wp=16383, bl=512, norm=True, clip=False)
    pack["rand_packed_bias"] = ISP.raw2bayer(raw, wp=16383, bl=512, norm=True, clip=True, bias=np.array([1, -2, 3, 0]))
    pack["rand_packed_nonorm"] = ISP.raw2bayer(raw, wp=16383, bl=512, norm=False)
    f = rng.rand(1, 4, 6, 8).astype(np.float32) * 1.2 - 0.1
    pack["unpack_in"] = f
    pack["unpack_out"] = ISP.bayer2raw(torch.from_numpy(f), wp=16383, bl=512)
    np.savez_compressed(os.path.join(OUT, "pack.npz"), **pack)

    # ---- S2/S3: parameter sampling known answers
    params = []
    for fn, cam, kw in (("sample_params", "SonyA7S2", {}), ("sample_params", "SonyA7S2", {"ln_ratio": True}),
                        ("sample_params", "CRVD", {}),
                        ("sample_params_max", "SonyA7S2", {}), ("sample_params_max", "SonyA7S2", {"iso": 1600}),
                        ("sample_params_max", "SonyA7S2", {"iso": 25600, "ratio": 200}),
                        ("sample_params_max", "IMX686", {}), ("sample_params_max", "IMX686", {"iso": 6400}),
                        ("sample_params_max", "IMX686", {"iso": 100, "ratio": 4}),
                        ("sample_params_max", "NikonD850", {}), ("sample_params_max", "CRVD", {})):
        for seed in (0, 1, 1997):
            np.random.seed(seed)
            p = getattr(P, fn)(cam, **kw)
            params.append({"fn": fn, "camera": cam, "kwargs": kw, "seed": seed, "out": _jsonable(p)})
    meta["params"] = params

    # ---- N1-N3: generate_noisy_obs with captured draws
    rng = np.random.RandomState(3)
    y = (rng.rand(4, 16, 24).astype(np.float32)) ** 2
    noisy = {"y": y}
    cases = []
    idx = 0
    for code in ("p", "pg", "pgr", "pgrq", "prq", "g", "grq", "pgrqd", "pb", "q", "", "r"):
        for chain in ("f64", "weak", "weak686", "f64r"):
            np.random.seed(11)
            if chain == "f64":
                p = P.sample_params("SonyA7S2")
            elif chain == "weak":
                p = P.sample_params_max("SonyA7S2", iso=3200)
            elif chain == "weak686":
                p = P.sample_params_max("IMX686", iso=6400)
            else:
                p = P.sample_params("SonyA7S2")
                p["ratio"] = np.float64(p["ratio"])
            if "d" in code and not isinstance(p["bias"], (np.ndarray, np.generic)):
                continue
            for ori, clip in ((False, False), (True, False), (False, True)):
                np.random.seed(100 + idx)
                z = P.generate_noisy_obs(y, param=p, noise_code=code, ori=ori, clip=clip)
                np.random.seed(100 + idx)          # re-draw in the reference's order to capture the draws
                pp = dict(p)
                pp["_lam"] = O.lam_of(y, p)[1]
                d = O.draw_reference_order(y.shape, pp, code)
                tag = f"c{idx}"
                noisy[tag + "_z"] = z
                for k, v in d.items():
                    noisy[f"{tag}_{k}"] = v
                cases.append({"tag": tag, "code": code, "chain": chain, "ori": ori, "clip": clip, "param": _jsonable(p)})
                idx += 1
    meta["noisy_obs_cases"] = cases
    np.savez_compressed(os.path.join(OUT, "noisy_obs.npz"), **noisy)

    # ---- N4: generate_noisy_torch with captured draws
    yt = torch.from_numpy(y)
    tz = {"y": y}
    tcases = []
    idx = 0
    for cam in ("SonyA7S2", "IMX686"):
        for code in ("p", "pr", "prq"):
            for ori, clip in ((False, False), (True, False), (False, 2)):
                np.random.seed(5 + idx)
                p = P.sample_params_max(cam)
                tp = {k: torch.from_numpy(np.array(v, np.float32)) for k, v in p.items()}
                torch.manual_seed(9 + idx)
                ref = P.generate_noisy_torch(yt.clone(), param=tp, noise_code=code, ori=ori, clip=clip).numpy()
                torch.manual_seed(9 + idx)
                yy = yt * (tp["wp"] - tp["bl"])
                yy = yy / tp["ratio"]
                tag = f"t{idx}"
                tz[tag + "_counts"] = torch.poisson(1.0 * yy / tp["K"]).numpy()
                tz[tag + "_read"] = torch.normal(torch.zeros_like(yy), (tp["sigGs"] / 1.0).expand(yy.shape)).numpy()
                if "r" in code:
                    tz[tag + "_row_z"] = torch.randn(4, 16, 1).numpy()
                if "q" in code:
                    tz[tag + "_q_u"] = torch.rand(yy.shape).numpy()
                tz[tag + "_z"] = ref
                tcases.append({"tag": tag, "camera": cam, "code": code, "ori": ori, "clip": clip, "param": _jsonable(p)})
                idx += 1
    meta["noisy_torch_cases"] = tcases
    np.savez_compressed(os.path.join(OUT, "noisy_torch.npz"), **tz)

    # ---- U1-U3: small networks (nf=4 keeps the fixture < 1 MB), reference init
    nets = {}
    arch = dict(name="UNetSeeInDark", in_nc=4, out_nc=4, nf=4, nframes=1, use_dpsv=False, res=False,
                cascade=False, add=False, lock_wb=False)
    x = torch.rand(2, 4, 32, 48, generator=torch.Generator().manual_seed(1997))
    nets["x"] = x.numpy()
    for name, res in (("UNetSeeInDark", False), ("UNetSeeInDark", True), ("ResUnet", False)):
        a = dict(arch, res=res)
        torch.manual_seed(7)
        net = getattr(R.archs, name)(a)
        R.archs.initialize_weights(net)
        net.eval()
        with torch.no_grad():
            out = net(x)
        tag = f"{name}_res{int(res)}"
        nets[tag + "__out"] = out.numpy()
        for k, v in net.state_dict().items():
            nets[f"{tag}__sd__{k}"] = v.numpy()
    np.savez_compressed(os.path.join(OUT, "nets.npz"), **nets)

    # ---- E1: illuminance correction + tensor2im
    ev = {}
    g = torch.Generator().manual_seed(3)
    pr = torch.rand(1, 4, 16, 20, generator=g) * 1.2 - 0.1
    src = torch.rand(1, 4, 16, 20, generator=g)
    src[0, 0, 0, :5] = 1
    ev["pred"], ev["src"] = pr.numpy(), src.numpy()
    ev["corrected"] = R.data_process.IlluminanceCorrect()(pr, src).numpy()
    ev["tensor2im"] = sys.modules["utils.visualization"].tensor2im(pr)
    np.savez_compressed(os.path.join(OUT, "eval.npz"), **ev)

    with open(os.path.join(OUT, "meta.json"), "w") as fh:
        json.dump(meta, fh, indent=1, sort_keys=True)
    for fn in sorted(os.listdir(OUT)):
        print(fn, os.path.getsize(os.path.join(OUT, fn)))


if __name__ == "__main__":
    main()
