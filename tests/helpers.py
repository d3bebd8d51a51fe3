"""Shared test helpers: golden-fixture loading and error metrics."""
import json
import os

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")

with open(os.path.join(GOLDEN, "index.json")) as _f:
    CASES = json.load(_f)["cases"]
CASE_NAMES = [c["name"] for c in CASES]
# cases of the path the oracle (oracle/reference_port.py, oracle/kernel_spec.py) restates: IRT link,
# Bernoulli responses.  The nonlinear decoders / Gaussian responses are pinned by
# the live-reference fixtures directly.
PORT_CASE_NAMES = [c["name"] for c in CASES if "generative" not in c]
CASE_BY_NAME = {c["name"]: c for c in CASES}


def load_case(name):
    cfg = CASE_BY_NAME[name]
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    rec = {k: z[k] for k in z.files}
    params = {k[len("param/"):]: torch.from_numpy(v) for k, v in rec.items() if k.startswith("param/")}
    grads = {k[len("grad/"):]: v for k, v in rec.items() if k.startswith("grad/")}
    return cfg, rec, params, grads


def rel_l2(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    den = np.linalg.norm(b)
    if den == 0:
        return float(np.linalg.norm(a))
    return float(np.linalg.norm(a - b) / den)


def max_rel(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    den = np.abs(b).max()
    if den == 0:
        return float(np.abs(a).max())
    return float(np.abs(a - b).max() / den)


def build_model(cfg, params, device="cpu"):
    """Drop-in module with the fixture's parameters loaded."""
    import vibo_b200
    cls = {1: vibo_b200.VIBO_1PL, 2: vibo_b200.VIBO_2PL, 3: vibo_b200.VIBO_3PL}[cfg["irt_model"]]
    model = cls(cfg["ability_dim"], cfg["I"], hidden_dim=64, ability_merge=cfg.get("merge", "product"),
                conditional_posterior=cfg["conditional"], generative_model=cfg.get("generative", "irt"),
                response_dist=cfg.get("response_dist", "bernoulli"),
                replace_missing_with_prior=not cfg["drop_missing"], n_norm_flows=cfg["n_flows"])
    missing, unexpected = model.load_state_dict(params, strict=True)
    return model.to(device)


def case_inputs(rec, device="cpu"):
    response = torch.from_numpy(rec["response"]).unsqueeze(2).to(device)
    mask = torch.from_numpy(rec["mask"]).bool().unsqueeze(2).to(device)
    eps_item = torch.from_numpy(rec["eps_item"]).to(device)
    eps_ability = torch.from_numpy(rec["eps_ability"]).to(device)
    return response, mask, eps_item, eps_ability
