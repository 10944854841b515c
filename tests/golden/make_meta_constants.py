"""Extracts the numeric constants and op attributes the reference's SHIPPED TensorFlow graphs hold for the hot path
 into tests/golden/meta_constants.json.

    python tests/golden/make_meta_constants.py       (build container only: reads /root/reference/**/*.meta)

The .meta files are serialized MetaGraphDefs written by TF 1.13.1 itself, so the values below are reference-held:
they are what the reference's arithmetic actually used (float32-rounded Python literals).  tests/test_oracle.py pins
the oracle's and the product's constants to this file, and (where /root/reference exists) this file to the graphs."""
import json
import os
import sys

import numpy as np
from tensorboard.compat.proto import meta_graph_pb2
from tensorboard.util import tensor_util

REF = "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "meta_constants.json")

PPO_META = "models/pretrained_agent/checkpoints/model.ckpt-705.meta"
VAE_META = "vae/models/rgb_bce_cnn_zdim64_beta1_kl_tolerance0.0_data/checkpoints/model.ckpt-232.meta"

PPO_CONSTS = {
    "log_prob_const": "policy/Normal/log_prob/add/x", "log_prob_half": "policy/Normal/log_prob/mul/x",
    "entropy_const": "policy/Normal/entropy/add/x", "clip_low": "clip_by_value/y", "clip_high": "clip_by_value/Minimum/y",
    "value_scale": "mul_2/y", "entropy_scale": "mul_3/y", "learning_rate": "ExponentialDecay/learning_rate",
    "adam_beta1": "Adam/beta1", "adam_beta2": "Adam/beta2", "adam_epsilon": "Adam/epsilon",
    "beta1_power_init": "beta1_power/initial_value", "beta2_power_init": "beta2_power/initial_value",
    "mean_affine_add": "policy/add/y", "mean_affine_div": "policy/truediv/y"}
VAE_CONSTS = {
    "range_low": "vae/GreaterEqual/y", "range_high": "vae/LessEqual/y", "reparam_half": "vae/mul/x",
    "kl_one": "vae/kl_divergence/add/x", "kl_minus_half": "vae/kl_divergence/mul/x", "beta": "vae/mul_1/x",
    "learning_rate": "vae/Adam/learning_rate", "adam_beta1": "vae/Adam/beta1", "adam_beta2": "vae/Adam/beta2",
    "adam_epsilon": "vae/Adam/epsilon", "beta1_power_init": "vae/beta1_power/initial_value",
    "beta2_power_init": "vae/beta2_power/initial_value"}


def load(path):
    m = meta_graph_pb2.MetaGraphDef()
    with open(os.path.join(REF, path), "rb") as f:
        m.ParseFromString(f.read())
    return m


def extract():
    out = {"source": "MetaGraphDefs shipped with the reference (written by TensorFlow %s)", "ppo": {}, "vae": {}}
    for key, path, table in (("ppo", PPO_META, PPO_CONSTS), ("vae", VAE_META, VAE_CONSTS)):
        m = load(path)
        nodes = {n.name: n for n in m.graph_def.node}
        out[key]["meta"] = path
        out[key]["tensorflow_version"] = m.meta_info_def.tensorflow_version
        out[key]["constants"] = {k: float(tensor_util.make_ndarray(nodes[v].attr["value"].tensor).reshape(-1)[0]) for k, v in table.items()}
        out[key]["variables"] = {n.name: [int(d.size) for d in n.attr["shape"].shape.dim]
                                 for n in m.graph_def.node if n.op == "VariableV2" and "Adam" not in n.name and "power" not in n.name}
    vae_nodes = load(VAE_META).graph_def.node
    out["vae"]["conv_ops"] = {n.name: {"op": n.op, "strides": list(n.attr["strides"].list.i), "padding": n.attr["padding"].s.decode(),
                                       "data_format": n.attr["data_format"].s.decode(), "dilations": list(n.attr["dilations"].list.i)}
                              for n in vae_nodes if n.op in ("Conv2D", "Conv2DBackpropInput") and "gradients" not in n.name}
    out["vae"]["gradient_ops"] = sorted({n.op for n in vae_nodes if "gradients" in n.name})
    ppo_nodes = {n.name: n for n in load(PPO_META).graph_def.node}
    # structure of the clipped surrogate: Minimum(ratio*A, clip(ratio)*A); the gradient's tie rule is LessEqual
    out["ppo"]["surrogate"] = {"Minimum_inputs": list(ppo_nodes["Minimum"].input),
                               "min_grad_select": sorted({n.op for n in ppo_nodes.values() if n.name.startswith("gradients/Minimum_grad/") and n.op in ("LessEqual", "Less", "Select")})}
    out["source"] = out["source"] % out["ppo"]["tensorflow_version"]
    return out


if __name__ == "__main__":
    blob = extract()
    with open(OUT, "w") as f:
        json.dump(blob, f, indent=1, sort_keys=True)
    print("wrote", OUT)
    print(json.dumps(blob["ppo"]["constants"], indent=1))
    print(json.dumps(blob["ppo"]["surrogate"], indent=1))
