"""Checkpoint wire format (reference purejaxql/utils/save_load.py:9-16,
pqn_minatar.py:467-483): safetensors files whose keys are the flax parameter paths joined with ",", shapes as flax
lays them out (conv kernel HWIO, dense kernels [in, out], NHWC flatten order for the 1024-row dense layer).
CPU only: the parameter layout comes from the C ABI (`pqn_net_layout`, host code), no kernel is launched."""
import numpy as np
import torch

MINATAR_TREE = {  # SURVEY Appendix C, Breakout (C=4, A=3)
    "BatchNorm_0,scale": (4,), "BatchNorm_0,bias": (4,),
    "CNN_0,Conv_0,kernel": (3, 3, 4, 16), "CNN_0,Conv_0,bias": (16,),
    "CNN_0,LayerNorm_0,scale": (16,), "CNN_0,LayerNorm_0,bias": (16,),
    "CNN_0,Dense_0,kernel": (1024, 128), "CNN_0,Dense_0,bias": (128,),
    "CNN_0,LayerNorm_1,scale": (128,), "CNN_0,LayerNorm_1,bias": (128,),
    "Dense_0,kernel": (128, 3), "Dense_0,bias": (3,),
}
CARTPOLE_TREE = {  # obs 4, A=2, HIDDEN_SIZE=256, NUM_LAYERS=2
    "BatchNorm_0,scale": (4,), "BatchNorm_0,bias": (4,),
    "Dense_0,kernel": (4, 256), "Dense_0,bias": (256,), "LayerNorm_0,scale": (256,), "LayerNorm_0,bias": (256,),
    "Dense_1,kernel": (256, 256), "Dense_1,bias": (256,), "LayerNorm_1,scale": (256,), "LayerNorm_1,bias": (256,),
    "Dense_2,kernel": (256, 2), "Dense_2,bias": (2,),
}


def _spec(kind, *a):
    from purejaxql_b200.networks import QNetworkSpec
    return QNetworkSpec(kind, *a)


def _check(spec, expected, tmp_path):
    from purejaxql_b200.utils.save_load import load_params, save_params, _flatten
    assert dict(zip(spec.flat_names(","), (tuple(e[2]) for e in spec.entries))) == expected
    # every tensor starts 16-byte aligned and the blocks do not overlap
    spans = sorted((int(off), int(off) + int(np.prod(shape))) for _, off, shape, _ in spec.entries)
    assert all(a % 4 == 0 for a, _ in spans)
    assert all(spans[i][1] <= spans[i + 1][0] for i in range(len(spans) - 1)) and spans[-1][1] <= spec.total
    S = 3
    flat = torch.arange(S * spec.total, dtype=torch.float32).reshape(S, spec.total)
    tree = spec.unflatten(flat)
    for s in range(S):  # one file per seed, as pqn_minatar.py:477-483 writes them
        one = {k: v[s] for k, v in _flatten(tree).items()}
        save_params(_nest(one), tmp_path / f"seed{s}.safetensors")
    back = [load_params(tmp_path / f"seed{s}.safetensors") for s in range(S)]
    for s in range(S):
        got = _flatten(back[s])
        assert {k: tuple(v.shape) for k, v in got.items()} == expected
        again = spec.flatten(back[s], device="cpu")
        assert torch.equal(again[0], _masked(spec, flat[s]))


def _nest(flatdict):
    tree = {}
    for k, v in flatdict.items():
        d = tree
        parts = k.split(",")
        for p in parts[:-1]:
            d = d.setdefault(p, {})
        d[parts[-1]] = v
    return tree


def _masked(spec, row):
    """The flat block may hold alignment padding between tensors; only the tensors travel through a checkpoint."""
    out = torch.zeros_like(row)
    for _, off, shape, _ in spec.entries:
        n = int(np.prod(shape))
        out[off:off + n] = row[off:off + n]
    return out


def test_minatar_checkpoint_names_shapes_and_roundtrip(tmp_path):
    from purejaxql_b200.networks import NET_CNN
    _check(_spec(NET_CNN, 4, 3), MINATAR_TREE, tmp_path)


def test_gymnax_checkpoint_names_shapes_and_roundtrip(tmp_path):
    from purejaxql_b200.networks import NET_MLP
    _check(_spec(NET_MLP, 4, 2, 256, 2), CARTPOLE_TREE, tmp_path)


def test_dense_kernel_rows_follow_nhwc_flatten_order():
    """Row (h*8+w)*16+c of CNN_0/Dense_0/kernel multiplies channel c of output pixel (h, w) (pqn_minatar.py:47):
    the conv kernels write h1 in exactly that order, so the [1024,128] view of the flat block needs no permutation."""
    from purejaxql_b200.networks import NET_CNN
    spec = _spec(NET_CNN, 4, 3)
    flat = torch.zeros((1, spec.total))
    tree = spec.unflatten(flat)
    w = tree["CNN_0"]["Dense_0"]["kernel"]
    assert w.shape == (1, 1024, 128) and w.stride()[1:] == (128, 1)
    k = tree["CNN_0"]["Conv_0"]["kernel"]
    assert k.shape == (1, 3, 3, 4, 16) and k.stride()[1:] == (3 * 4 * 16, 4 * 16, 16, 1)  # HWIO, O fastest
