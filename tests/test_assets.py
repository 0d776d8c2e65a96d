"""CPU: asset ingestion that needs no GPU -- the SMPL pickle loader without chumpy."""
import pickle
import sys
import types

import numpy as np
import scipy.sparse as sp


def _fake_chumpy_pickle(model, path):
    """Write a pickle that looks like the official SMPL file: chumpy.ch.Ch-wrapped arrays + scipy sparse regressors."""
    mod = types.ModuleType('chumpy'); sub = types.ModuleType('chumpy.ch')

    class Ch(object):
        def __init__(self, x):
            self.x = np.asarray(x)
            self.dterms = ()
    Ch.__module__ = 'chumpy.ch'; Ch.__qualname__ = 'Ch'
    sub.Ch = Ch; mod.ch = sub
    sys.modules['chumpy'] = mod; sys.modules['chumpy.ch'] = sub
    try:
        dd = dict(model)
        for k in ('v_template', 'shapedirs', 'posedirs', 'weights'):
            dd[k] = Ch(model[k])
        dd['J_regressor'] = sp.csc_matrix(model['J_regressor'])
        dd['cocoplus_regressor'] = sp.csc_matrix(model['cocoplus_regressor'])
        with open(path, 'wb') as f:
            pickle.dump(dd, f, protocol=2)
    finally:
        del sys.modules['chumpy'], sys.modules['chumpy.ch']


def test_smpl_pickle_loads_without_chumpy(tmp_path, smpl_model):
    from human_dynamics_b200.smpl import load_smpl_model, _dense
    path = str(tmp_path / 'smpl.pkl')
    _fake_chumpy_pickle(smpl_model, path)
    assert 'chumpy' not in sys.modules
    dd = load_smpl_model(path)
    for k in ('v_template', 'shapedirs', 'posedirs', 'weights', 'J_regressor', 'cocoplus_regressor'):
        assert np.array_equal(_dense(dd[k]), np.asarray(smpl_model[k])), k
    assert np.array_equal(np.asarray(dd['kintree_table']), smpl_model['kintree_table'])
    # the oracle reads the same loaded dict (dense / sparse / chumpy-stub all accepted)
    from oracle.smpl_ref import SMPLRef
    from human_dynamics_b200 import synthetic
    beta, theta = synthetic.make_smpl_inputs(2, seed=1)
    dd_plain = {k: (_dense(v) if k in ('v_template', 'shapedirs', 'posedirs', 'weights') else v) for k, v in dd.items()}
    v1 = SMPLRef(dd_plain)(beta, theta, get_skin=True)[0]
    v0 = SMPLRef(smpl_model)(beta, theta, get_skin=True)[0]
    assert np.allclose(v0, v1, atol=1e-6)          # same model through dense vs sparse->dense regressors (BLAS path may differ)
    assert load_smpl_model(smpl_model) is smpl_model
