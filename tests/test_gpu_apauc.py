"""GPU suite: the UNMODIFIED reference scripts (``oracle/_ref/TPNet``, installed by ``oracle/make_ref.py``)
on the drop-in — north_star's third parity criterion, AP/AUC within 0.1 points.

Down-scaled Wikipedia-shaped dataset (bipartite, 2 projection layers, batch 200, K=20), 2 epochs.
``scripts/apauc_parity.py`` runs ``train_link_prediction.py`` with the reference's own class and with
``tpnet_b200.launch`` swapping in the CUDA class, then evaluates the stock arm's checkpoint with
``evaluate_link_prediction.py`` under both classes.  Full-shape results: ``profiles/r02_apauc_*.json``.
"""
import os
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'scripts'))

TOL = 0.001          # 0.1 points of AP / AUC (north_star)


@pytest.fixture(scope='module')
def parity():
    import apauc_parity
    if not os.path.isfile(os.path.join(apauc_parity.REF, 'models', 'TPNet.py')):
        pytest.skip('oracle/_ref/TPNet not installed (python oracle/make_ref.py needs /root/reference)')
    return apauc_parity.parity_run('wikipedia', epochs=2, edges=12000, src=400, dst=60, timeout=1500)


def test_same_checkpoint_same_ap_auc_under_both_classes(parity):
    """One checkpoint, evaluate_link_prediction.py under the reference class and under the drop-in: test and
    new-node test AP / AUC agree within 0.1 points (same weights and negatives: the hot path is the only difference)."""
    assert set(parity['eval_stock']) == {'test metrics', 'new node test metrics'}
    for split, metrics in parity['eval_stock'].items():
        for name, value in metrics.items():
            assert abs(value - parity['eval_dropin'][split][name]) <= TOL, (split, name, parity)
    assert all(0.0 < v <= 1.0 for m in parity['eval_dropin'].values() for v in m.values())


def test_training_on_the_drop_in_reaches_the_reference_ap_auc(parity):
    """train_link_prediction.py end to end on each class (reset per epoch, update per batch, backup / reload around
    validation, early-stopping checkpoints through state_dict): validate / test / new-node AP and AUC within 0.1 points."""
    for split in ('validate metrics', 'new node validate metrics', 'test metrics', 'new node test metrics'):
        for name, value in parity['train_stock'][split].items():
            assert abs(value - parity['train_dropin'][split][name]) <= TOL, (split, name, parity)
