"""CPU checks of the reference arm's plumbing (baseline/): the staged copy of the reference's Python is byte-identical to its source,
imports unmodified in this environment, builds the 14.13 M-parameter backbone the survey counted, and importing the arm does not pull
the product (epnet_b200 / libepnet_b200.so) or the oracle package into the process."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _staged():
    from baseline import stage_ref
    if not stage_ref.staged():
        if not os.path.isdir(stage_ref.SRC):
            pytest.skip("baseline/_ref not staged and /root/reference absent")
        stage_ref.stage()
    return stage_ref


def test_staged_files_are_byte_identical_to_the_reference():
    stage_ref = _staged()
    if not os.path.isdir(stage_ref.SRC):
        pytest.skip("/root/reference absent (GPU box): nothing to compare against")
    assert stage_ref.verify() == []
    assert os.path.exists(os.path.join(stage_ref.DEST, "lib", "net", "pointnet2_msg.py"))
    assert os.path.exists(os.path.join(stage_ref.DEST, "tools", "cfgs", "LI_Fusion_with_attention_use_ce_loss.yaml"))


def test_reference_arm_process_does_not_load_the_product():
    _staged()
    code = ("import sys; sys.path.insert(0, %r); from baseline import ref_arm, ref_env; import torch\n"
            "ref = ref_env.import_reference('reference')\n"
            "torch.manual_seed(0)\n"
            "net = ref.pointnet2_msg.Pointnet2MSG(input_channels=0, use_xyz=True)\n"
            "n = sum(p.numel() for p in net.parameters())\n"
            "bad = [m for m in sys.modules if m == 'epnet_b200' or m.startswith('epnet_b200.') or m == 'oracle' or m.startswith('oracle.')]\n"
            "maps = [l for l in open('/proc/self/maps') if 'libepnet_b200' in l or 'liboracle' in l]\n"
            "print('PARAMS', n, 'BAD', bad, 'MAPS', len(maps))\n" % ROOT)
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    line = [l for l in out.stdout.splitlines() if l.startswith("PARAMS")][-1]
    assert "PARAMS 14131949 BAD [] MAPS 0" == line, line  # 14,131,949 parameters with the yaml config


def test_reference_config_is_the_published_yaml():
    _staged()
    from baseline import ref_env
    ref = ref_env.import_reference("reference")
    cfg = ref.cfg
    assert list(cfg.RPN.SA_CONFIG.NPOINTS) == [4096, 1024, 256, 64] and cfg.LI_FUSION.ENABLED and cfg.LI_FUSION.ADD_Image_Attention
    assert cfg.RPN.USE_INTENSITY is False and list(cfg.LI_FUSION.DeConv_Kernels) == [2, 4, 8, 16]
