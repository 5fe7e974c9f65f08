"""the reference's UNMODIFIED entry point — main_instance_segmentation.py (:21-118) with its conf/ tree,
trainer/trainer.py (InstanceSegmentation.__init__ :46-91, training_step :99-163, configure_optimizers :953-966, train_dataloader
:976-982), datasets/utils.py (FreeMaskVoxelizeCollate), models/*.py — runs training steps on the CUDA shim.

hydra, omegaconf and pytorch_lightning are absent from this image: unscene3d_b200/standins provides minimal stand-ins (hydra 1.0
composition and instantiate, attribute-style configs with interpolation, a plain fit loop); a real installation of
any of them takes precedence.  The run is launched the way scripts/unsupervised/train_unscene3d.sh launches it — config groups
and values overridden on the command line, nothing edited — with the dataset node pointed at a synthetic dataset that yields
samples in the reference dataset's layout.
"""
import os
import runpy
import sys

import pytest
import torch

from helpers import staged_reference_root

pytestmark = [pytest.mark.skipif(staged_reference_root() is None, reason="no reference tree (neither /root/reference nor oracle/_ref/reference)")]

OVERRIDES = [
    "general.experiment_name=us3d_entry_point", "general.project_name=unscene3d", "general.eval_on_segments=true",
    "general.train_on_segments=true", "general.num_targets=3", "data.batch_size=2", "data.test_batch_size=1", "data.num_workers=0",
    "data/collation_functions=freemask_voxelize_collate", "data/datasets=freemask", "general.resume=False",
    "data.train_dataset._target_=unscene3d_b200.synthetic.SyntheticFreemaskDataset",
    "data.validation_dataset._target_=unscene3d_b200.synthetic.SyntheticFreemaskDataset",
    "data.test_dataset._target_=unscene3d_b200.synthetic.SyntheticFreemaskDataset",
    "trainer.max_epochs=2", "+trainer.limit_train_batches=2",
]


def test_conf_tree_composes_like_the_training_script_asks():
    """CPU: the stand-in hydra composes the unmodified conf/ tree with the overrides of scripts/unsupervised/train_unscene3d.sh."""
    from unscene3d_b200 import standins

    standins.install()
    import hydra

    if "us3d-standin" not in getattr(hydra, "__version__", ""):
        pytest.skip("a real hydra is installed")
    cfg = hydra.compose(os.path.join(staged_reference_root(), "conf"), "config_base_instance_segmentation.yaml", OVERRIDES)
    assert cfg.model._target_ == "models.mask3d.Mask3D" and cfg.model.num_classes == 3 and cfg.model.train_on_segments is True
    assert cfg.model.config.backbone._target_ == "models.res16unet.Res16UNet34C" and cfg.model.config.backbone.in_channels == 3
    assert cfg.data.train_collation._target_ == "datasets.utils.FreeMaskVoxelizeCollate" and cfg.data.train_collation.voxel_size == 0.02
    assert cfg.data.train_dataset._target_ == "unscene3d_b200.synthetic.SyntheticFreemaskDataset" and cfg.data.train_dataset.mode == "train"
    assert cfg.matcher.cost_mask == 5.0 and cfg.loss.num_points == -1 and cfg.optimizer.lr == 1e-4
    assert cfg.scheduler.scheduler.max_lr == 1e-4 and cfg.scheduler.scheduler.epochs == 2 and cfg.trainer.limit_train_batches == 2
    assert cfg.general.save_dir == "saved/us3d_entry_point" and [c._target_ for c in cfg.callbacks][0] == "pytorch_lightning.callbacks.ModelCheckpoint"


@pytest.mark.gpu
def test_unmodified_entry_point_runs_training_steps_on_the_shim(tmp_path):
    import unscene3d_b200  # noqa: F401  (shims first on sys.path)
    from unscene3d_b200 import _lib, standins

    standins.install()
    import pytorch_lightning as pl

    if "us3d-standin" not in getattr(pl, "__version__", ""):
        pytest.skip("a real pytorch_lightning is installed: the stand-in loop is not in use")
    root = staged_reference_root()
    for k in [k for k in sys.modules if k == "models" or k.startswith(("models.", "trainer", "datasets", "utils.", "benchmark")) or k == "utils"]:
        del sys.modules[k]
    captured = {}
    orig_fit = pl.Trainer.fit

    def fit(self, model):
        captured["trainer"], captured["model"] = self, model
        captured["before"] = {k: v.detach().clone() for k, v in list(model.named_parameters())[:3]}
        return orig_fit(self, model)

    pl.Trainer.fit = fit
    argv, cwd = sys.argv, os.getcwd()
    sys.path.insert(0, root)
    os.chdir(tmp_path)
    _lib.reset_launch_count()
    try:
        sys.argv = ["main_instance_segmentation.py"] + OVERRIDES + [f"general.save_dir={tmp_path}/saved", "general.gpus=1"]
        runpy.run_path(os.path.join(root, "main_instance_segmentation.py"), run_name="__main__")
    finally:
        sys.argv = argv
        os.chdir(cwd)
        sys.path.remove(root)
        pl.Trainer.fit = orig_fit
    trainer, model = captured["trainer"], captured["model"]
    assert root in sys.modules["trainer.trainer"].__file__ and root in sys.modules["models.mask3d"].__file__
    assert type(model).__name__ == "InstanceSegmentation" and type(model.model).__module__ == "models.mask3d"
    assert trainer.global_step == 4 and len(trainer.losses) == 4 and all(torch.isfinite(torch.tensor(trainer.losses)))
    assert _lib.launch_count() > 4 * 500, "the steps did not run on libus3d"
    changed = [float((p.detach().cpu() - captured["before"][k].cpu()).abs().max()) for k, p in list(model.named_parameters())[:3]]
    assert max(changed) > 0, "the optimizer did not update the model"
    assert "train_loss_mask" in model.logged and "train_mean_loss_dice" in model.logged
    assert os.path.isfile(f"{tmp_path}/saved/last-epoch.ckpt"), "RegularCheckpointing did not fire"
    print(f"entry point: 4 training steps, losses {[round(l, 3) for l in trainer.losses]}, {_lib.launch_count()} libus3d launches")


def _run_entry_point(root, tmp_path, overrides, on_fit):
    """Runs the unmodified main_instance_segmentation.py with `overrides`; `on_fit(trainer, model)` sees the model right before fit."""
    from unscene3d_b200 import standins

    standins.install()
    import pytorch_lightning as pl

    for k in [k for k in sys.modules if k == "models" or k.startswith(("models.", "trainer", "datasets", "utils.", "benchmark")) or k == "utils"]:
        del sys.modules[k]
    orig_fit = pl.Trainer.fit

    def fit(self, model):
        on_fit(self, model)
        return orig_fit(self, model)

    pl.Trainer.fit = fit
    argv, cwd = sys.argv, os.getcwd()
    sys.path.insert(0, root)
    os.chdir(tmp_path)
    try:
        sys.argv = ["main_instance_segmentation.py"] + overrides
        runpy.run_path(os.path.join(root, "main_instance_segmentation.py"), run_name="__main__")
    finally:
        sys.argv = argv
        os.chdir(cwd)
        sys.path.remove(root)
        pl.Trainer.fit = orig_fit


@pytest.mark.gpu
def test_unmodified_entry_point_loads_checkpoints_on_the_shim(tmp_path):
    """SURVEY §8(f1), checkpoint half: a checkpoint written by one run of the unmodified entry point (RegularCheckpointing,
    trainer/trainer.py:30-36) is loaded by the next through the reference's own loaders — `general.checkpoint` ->
    load_checkpoint_with_missing_or_exsessive_keys (utils/utils.py:98-128) and `general.backbone_checkpoint` ->
    load_backbone_checkpoint_with_missing_or_exsessive_keys (utils/utils.py:58-96): every parameter and buffer name of the shim's
    modules is the reference's, so nothing is reported missing and the loaded model equals the saved one."""
    import unscene3d_b200  # noqa: F401
    from unscene3d_b200 import standins

    standins.install()
    import pytorch_lightning as pl

    if "us3d-standin" not in getattr(pl, "__version__", ""):
        pytest.skip("a real pytorch_lightning is installed: the stand-in loop is not in use")
    root = staged_reference_root()
    base = [o for o in OVERRIDES if not o.startswith(("trainer.max_epochs", "+trainer.limit_train_batches"))]
    base += ["trainer.max_epochs=1", "+trainer.limit_train_batches=1", "general.gpus=1"]
    # run 1: one step, writes <save_dir>/last-epoch.ckpt
    _run_entry_point(root, tmp_path, base + [f"general.save_dir={tmp_path}/run1"], lambda tr, m: None)
    ckpt = f"{tmp_path}/run1/last-epoch.ckpt"
    saved = torch.load(ckpt)["state_dict"]
    assert any(k.startswith("model.backbone.") for k in saved) and any(k.startswith("model.mask_embed_head") for k in saved)
    # run 2: general.checkpoint -> the whole model comes from the file
    seen = {}
    _run_entry_point(root, tmp_path, base + [f"general.save_dir={tmp_path}/run2", f"general.checkpoint={ckpt}"],
                     lambda tr, m: seen.update({k: v.detach().cpu().clone() for k, v in m.state_dict().items()}))
    assert set(seen) == set(saved)
    assert all(torch.equal(seen[k], saved[k].cpu()) for k in saved), "the loaded model differs from the checkpoint"
    # run 3: general.backbone_checkpoint -> a backbone-only file (keys without the `model.backbone.` prefix)
    bb = {k[len("model.backbone."):]: v for k, v in saved.items() if k.startswith("model.backbone.")}
    bb_path = f"{tmp_path}/backbone.ckpt"
    torch.save({"state_dict": bb}, bb_path)
    seen3 = {}
    _run_entry_point(root, tmp_path, base + [f"general.save_dir={tmp_path}/run3", f"general.backbone_checkpoint={bb_path}"],
                     lambda tr, m: seen3.update({k: v.detach().cpu().clone() for k, v in m.state_dict().items()}))
    assert all(torch.equal(seen3["model.backbone." + k], v.cpu()) for k, v in bb.items()), "the backbone differs from the checkpoint"
    head = [k for k in saved if k.startswith("model.mask_embed_head") and saved[k].dtype.is_floating_point]
    assert any(not torch.equal(seen3[k], saved[k].cpu()) for k in head), "the decoder should have kept its fresh initialisation"
