"""Helpers to read the committed reference fixtures (tests/golden/*.npz)."""
import os

import numpy as np

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')
CASES = ['wiki_tiny', 'flights_tiny', 'reddit_tiny', 'one_layer', 'four_layer', 'matrix_kat']


def load_case(name):
    z = np.load(os.path.join(GOLDEN_DIR, name + '.npz'))
    cfg = {k[4:]: z[k].item() for k in z.files if k.startswith('cfg_')}
    nb = int(z['n_batches'])
    batches = [(z[f'b{b}_src'], z[f'b{b}_dst'], z[f'b{b}_t'], z[f'b{b}_w']) for b in range(nb)]
    return z, cfg, batches


def oracle_kwargs(cfg):
    return dict(node_num=int(cfg['node_num']), edge_num=int(cfg['edge_num']), dim_factor=int(cfg['dim_factor']),
                num_layer=int(cfg['num_layer']), time_decay_weight=float(cfg['time_decay_weight']),
                use_matrix=bool(cfg['use_matrix']), beginning_time=float(cfg['beginning_time']),
                not_scale=bool(cfg['not_scale']), enforce_dim=int(cfg['enforce_dim']))
