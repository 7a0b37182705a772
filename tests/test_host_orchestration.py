"""CPU: the product's host-side orchestration (edgecape_b200/*.py: views, strides, argument
order, buffer reuse, control flow, result assembly) driven end to end through the reference-facing
API with the C ABI *emulated* on host pointers (tests/cpu_emulator.py), against the golden vectors
frozen from the unmodified reference.  This does not test the CUDA kernels (tests -m gpu do); it
makes sure a kernel-correct library is called correctly."""
import os

import numpy as np
import pytest
import torch

import edgecape_b200 as E
from edgecape_b200.config import state_dict_shapes
from edgecape_b200.synthetic import make_state_dict
from oracle.gen_golden import build_case

from . import cpu_emulator

CASES = ["c1_tiny", "tiny_k100_2shot_masked", "tiny_allmasked"]


def _np(t):
    return t.detach().float().cpu().numpy() if torch.is_tensor(t) else np.asarray(t)


@pytest.mark.parametrize("tensor_cores", [True, False, "fused_gcn", "fused_gcn_aggregate_first"],
                         ids=["split_f16_path", "fp32_path", "fused_gcn", "fused_gcn_aggregate_first"])
@pytest.mark.parametrize("name", CASES)
def test_host_orchestration_against_golden(name, tensor_cores, golden_dir, monkeypatch):
    from edgecape_b200 import ops
    cpu_emulator.install(monkeypatch)
    monkeypatch.setattr(ops, "GCN_FUSED", {"fused_gcn": 2, "fused_gcn_aggregate_first": 1}.get(tensor_cores, 0))
    tensor_cores = bool(tensor_cores)
    monkeypatch.setattr(ops, "TENSOR_CORES", tensor_cores)
    golden = dict(np.load(os.path.join(golden_dir, name + ".npz")))
    cfg, data, wseed = build_case(name)
    model = E.build_model(dict(model=cfg))
    model.load_state_dict(make_state_dict(state_dict_shapes(cfg), wseed), strict=True)
    model.eval()
    res = model(return_loss=False, **data)
    _, inter = model.predict(data["img_s"], data["target_s"], data["target_weight_s"], data["img_q"],
                             data["img_metas"], return_intermediates=True)
    feat_q, _ = model.extract_features(data["img_s"], data["img_q"])
    got = dict(inter)
    got.update(feature_q=feat_q, preds=res["preds"], boxes=res["boxes"], points=res["points"], skeleton=res["skeleton"])
    for k, w in golden.items():
        if k not in got or got[k] is None:
            continue
        a = _np(got[k])
        if k in ("feature_q", "encoder_image"):
            a = a[:1]
        if k == "argmax":
            assert np.array_equal(a, w), f"{name}:{k}"
            continue
        assert a.shape == w.shape, (k, a.shape, w.shape)
        err = np.abs(a.astype(np.float64) - w).max() / (np.abs(w).max() + 1e-12)
        assert err < 2e-4, f"{name}:{k} rel err {err:.3e}"
    assert res["image_paths"] == [m["query_image_file"] for m in data["img_metas"]]
    assert res["bbox_ids"] == [m["bbox_id"] for m in data["img_metas"]]


def test_support_deduplication_changes_nothing_but_the_backbone_batch(monkeypatch):
    """test_cfg['dedup_supports']: rows sharing `sample_image_file` run the backbone once per distinct support; the
    result dict must equal the reference behaviour (every row's own copy through the ViT)."""
    from edgecape_b200 import ops
    from edgecape_b200.synthetic import make_episode
    from oracle.gen_golden import TINY_VIT, model_cfg_for
    cpu_emulator.install(monkeypatch)
    monkeypatch.setattr(ops, "TENSOR_CORES", False)
    cfg = model_cfg_for(TINY_VIT)
    data = make_episode(batch=5, image_size=64, num_kpts=7, shots=2, seed=11, shared_support=3)
    model = E.build_model(dict(model=cfg))
    model.load_state_dict(make_state_dict(state_dict_shapes(cfg), 5), strict=True)
    model.eval()
    want = model(return_loss=False, **data)
    calls = []
    orig = model.encoder_query.forward_tokens
    monkeypatch.setattr(model.encoder_query, "forward_tokens",
                        lambda images: (calls.append([int(t.shape[0]) for t in images]), orig(images))[1])
    model.test_cfg = dict(model.test_cfg, dedup_supports=True)
    assert model._support_groups(data["img_metas"]) == ([0, 3], [0, 0, 0, 1, 1])
    got = model(return_loss=False, **data)
    assert calls == [[5, 2, 2]]                      # 5 queries + 2 distinct supports per shot instead of 5 + 5 + 5
    for k in ("preds", "points", "skeleton", "boxes"):
        assert np.allclose(np.asarray(got[k]), np.asarray(want[k]), rtol=0, atol=1e-6), k


def test_test_loop_api_returns_results_in_order(monkeypatch):
    """edgecape_b200.apis.single_gpu_test (the reference's apis/test.py loop): one result dict per batch, in order;
    without CUDA graphs it degrades to one synchronous forward_test per batch."""
    from edgecape_b200 import ops
    from edgecape_b200.apis import iter_results, single_gpu_test
    from edgecape_b200.synthetic import make_episode
    from oracle.gen_golden import TINY_VIT, model_cfg_for
    cpu_emulator.install(monkeypatch)
    monkeypatch.setattr(ops, "TENSOR_CORES", False)
    cfg = model_cfg_for(TINY_VIT)
    model = E.build_model(dict(model=cfg))
    model.load_state_dict(make_state_dict(state_dict_shapes(cfg), 5), strict=True)
    model.eval()
    model.use_cuda_graph = False
    batches = [make_episode(batch=2, image_size=64, num_kpts=5, shots=1, seed=20 + i) for i in range(3)]
    want = [model(return_loss=False, **d) for d in batches]
    got = single_gpu_test(model, batches)
    assert [g["image_paths"] for g in got] == [w["image_paths"] for w in want]
    for g, w in zip(got, want):
        assert np.array_equal(g["preds"], w["preds"])
    assert len(list(iter_results(model, iter(batches[:1]), depth=4))) == 1


class DataContainer:
    """Shape of mmcv.parallel.DataContainer after `collate` (mmcv is absent offline): `.data` is a list with one
    entry per GPU; stacked tensors for `stack=True` fields, the list of meta dicts for `cpu_only=True`."""

    def __init__(self, data, stack=False, cpu_only=False):
        self.data, self.stack, self.cpu_only = data, stack, cpu_only


def test_data_container_batches_are_unwrapped(monkeypatch):
    """The reference's loop hands the model what mmcv's collate produced; only MMDataParallel.scatter unwraps the
    DataContainers (apis/test.py:33).  INTEGRATION.md drops that wrapper, so forward() and apis.iter_results must
    accept the wrapped batch themselves."""
    from edgecape_b200 import ops
    from edgecape_b200.apis import single_gpu_test
    from edgecape_b200.synthetic import make_episode
    from oracle.gen_golden import TINY_VIT, model_cfg_for
    cpu_emulator.install(monkeypatch)
    monkeypatch.setattr(ops, "TENSOR_CORES", False)
    cfg = model_cfg_for(TINY_VIT)
    model = E.build_model(dict(model=cfg))
    model.load_state_dict(make_state_dict(state_dict_shapes(cfg), 5), strict=True)
    model.eval()
    model.use_cuda_graph = False
    data = make_episode(batch=2, image_size=64, num_kpts=5, shots=2, seed=31)
    want = model(return_loss=False, **data)
    wrapped = dict(data)
    wrapped["img_metas"] = DataContainer([data["img_metas"]], cpu_only=True)
    wrapped["img_q"] = DataContainer([data["img_q"]], stack=True)
    wrapped["img_s"] = [DataContainer([t], stack=True) for t in data["img_s"]]
    for got in (model(return_loss=False, **wrapped), single_gpu_test(model, [wrapped])[0]):
        assert got["image_paths"] == want["image_paths"]
        for k in ("preds", "points", "skeleton", "boxes"):
            assert np.array_equal(np.asarray(got[k]), np.asarray(want[k])), k


def test_support_deduplication_key_includes_the_crop():
    """Two rows with the same support image FILE but different annotated instances (another bbox / crop) must not
    share backbone features (MP-100 pairs are drawn per object id, test_dataset.py:93-97)."""
    from edgecape_b200.synthetic import make_episode
    from oracle.gen_golden import TINY_VIT, model_cfg_for
    model = E.build_model(dict(model=model_cfg_for(TINY_VIT)))
    model.test_cfg = dict(model.test_cfg, dedup_supports=True)
    metas = make_episode(batch=4, image_size=64, num_kpts=5, shots=1, seed=3, shared_support=4)["img_metas"]
    assert model._support_groups(metas) == ([0], [0, 0, 0, 0])
    metas[2]["sample_center"] = [np.array([10.0, 12.0], dtype=np.float32)]        # same file, another instance
    assert model._support_groups(metas) == ([0, 2], [0, 0, 1, 0])
    for m in metas:                                                             # no crop information: never merged
        for k in ("sample_center", "sample_scale", "sample_rotation"):
            m.pop(k)
    assert model._support_groups(metas) is None


def test_host_decode_equals_per_sample_transform_preds():
    """TwoStageHead.decode (head.py:324-387) on host arrays: the batched form equals transform_preds row by row."""
    from edgecape_b200.head import transform_preds
    from edgecape_b200.synthetic import make_episode
    from oracle.gen_golden import TINY_VIT, model_cfg_for
    model = E.build_model(dict(model=model_cfg_for(TINY_VIT)))
    head = model.keypoint_head_module
    metas = make_episode(batch=3, image_size=64, num_kpts=6, shots=1, seed=4)["img_metas"]
    rng = np.random.default_rng(0)
    for i, m in enumerate(metas):
        m["query_center"] = rng.uniform(20, 200, 2).astype(np.float32)
        m["query_scale"] = rng.uniform(0.3, 2.0, 2).astype(np.float32)
        m["query_bbox_score"] = 0.5 + 0.1 * i
    out = rng.uniform(0, 1, (3, 6, 2))
    for udp in (False, True):
        head.test_cfg = dict(head.test_cfg or {}, use_udp=udp)
        res = head.decode(metas, out, (64, 48))
        for b, m in enumerate(metas):
            want = transform_preds(out[b] * np.array([64, 48]), m["query_center"], m["query_scale"], [64, 48], use_udp=udp)
            assert np.allclose(res["preds"][b, :, :2], want, rtol=1e-6, atol=1e-4)
            assert np.all(res["preds"][b, :, 2] == 1.0)
            assert np.allclose(res["boxes"][b], [*m["query_center"], *m["query_scale"],
                                                 np.prod(m["query_scale"] * 200.0), 0.5 + 0.1 * b], rtol=1e-6)
    assert res["bbox_ids"] == [0, 1, 2] and res["image_paths"] == [m["query_image_file"] for m in metas]
