"""oracle/_ref, Python edition: the reference's own source EXECUTED, unmodified.

TEST INFRASTRUCTURE ONLY.  The reference (vshmyhlo/retinanet-tensorflow) is TensorFlow-1.x Python; TensorFlow is
not installable here, but the dense-head path uses ~60 TF ops, all elementwise / reduce / gather.  `tf_stub/`
provides those ops on NumPy with TF 1.x semantics (dtype conversion of operands, int true-division in float64,
linspace, argmax ties, one_hot, boolean_mask, huber_loss, non_max_suppression, tf.gradients with TF's registered
gradient formulas).  `load()` puts the stub and /root/reference on sys.path, imports the reference modules
`levels`, `utils`, `dataset`, `losses`, `augmentation` from where they lie -- no copy, no edit -- and returns them.

This pins op order, masks, axis choices and the loss formulas to the reference's TEXT instead of a restatement:
tests/golden/make_ref_fixtures.py runs these modules on seeded inputs and commits the outputs
(tests/golden/ref_exec_*.npz); tests/test_oracle_golden.py compares oracle/rn_oracle.py against them on CPU and
tests/test_gpu_golden.py compares the CUDA path against them on the GPU.  /root/reference exists only in the
build container, so only the fixture generator (and the CPU test that re-runs it when the reference is present)
call load().
"""
import importlib
import os
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
STUB = os.path.join(HERE, 'tf_stub')
REFERENCE = os.environ.get('RN_REFERENCE_DIR', '/root/reference')
MODULES = ('levels', 'utils', 'augmentation', 'dataset', 'losses')
_SHADOWED = ('tensorflow', 'tensorflow.python', 'tensorflow.python.client', 'tensorflow.python.client.device_lib',
             'termcolor') + MODULES


def available():
    return all(os.path.exists(os.path.join(REFERENCE, m + '.py')) for m in MODULES)


def load():
    """Returns a namespace with .tf (the stub) and the reference modules .levels .utils .dataset .losses
    .augmentation, imported unmodified from REFERENCE.  sys.modules / sys.path are restored afterwards, so the
    bare module names (`utils`, `losses`, ...) do not leak into the caller's process."""
    if not available():
        raise FileNotFoundError('reference checkout not found at %s' % REFERENCE)
    saved = {name: sys.modules.pop(name) for name in _SHADOWED if name in sys.modules}
    sys.path[:0] = [STUB, REFERENCE]
    try:
        ns = types.SimpleNamespace()
        ns.tf = importlib.import_module('tensorflow')
        assert ns.tf.__version__.endswith('numpy-stub'), 'a real tensorflow shadowed the stub'
        for m in MODULES:
            setattr(ns, m, importlib.import_module(m))
        for m in MODULES:
            f = os.path.realpath(getattr(ns, m).__file__)
            assert f.startswith(os.path.realpath(REFERENCE) + os.sep), '%s was not imported from the reference: %s' % (m, f)
        return ns
    finally:
        for name in _SHADOWED:
            sys.modules.pop(name, None)
        sys.modules.update(saved)
        for p in (STUB, REFERENCE):
            if p in sys.path:
                sys.path.remove(p)


def run_reference_tests(verbosity=0):
    """Runs the reference's OWN unittest files for this path (utils_test.py, losses_test.py, dataset_test.py,
    augmentation_test.py) under the stub.  Returns {test id: 'ok' | 'fail: ...' | 'error: ...'}.  Some of those
    tests are stale in the reference itself (they call functions with signatures the code no longer has); the
    caller decides which ones must pass (tests/test_ref_exec.py lists them)."""
    import unittest
    saved = {name: sys.modules.pop(name) for name in _SHADOWED if name in sys.modules}
    sys.path[:0] = [STUB, REFERENCE]
    results = {}
    try:
        for mod_name in ('utils_test', 'losses_test', 'dataset_test', 'augmentation_test'):
            sys.modules.pop(mod_name, None)
            mod = importlib.import_module(mod_name)
            suite = unittest.defaultTestLoader.loadTestsFromModule(mod)

            def walk(s):
                for t in s:
                    if isinstance(t, unittest.TestSuite):
                        yield from walk(t)
                    else:
                        yield t
            ids = [t.id() for t in walk(suite)]
            res = unittest.TextTestRunner(stream=open(os.devnull, 'w'), verbosity=verbosity).run(suite)
            bad = {t.id(): 'fail: ' + tb.strip().splitlines()[-1] for t, tb in res.failures}
            bad.update({t.id(): 'error: ' + tb.strip().splitlines()[-1] for t, tb in res.errors})
            for i in ids:
                results[i] = bad.get(i, 'ok')
            sys.modules.pop(mod_name, None)
        return results
    finally:
        for name in _SHADOWED:
            sys.modules.pop(name, None)
        sys.modules.update(saved)
        for p in (STUB, REFERENCE):
            if p in sys.path:
                sys.path.remove(p)


if __name__ == '__main__':
    for k, v in sorted(run_reference_tests().items()):
        print('%-70s %s' % (k, v))
