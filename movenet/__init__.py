"""Import shim: ``from movenet.wavenet import WaveNet`` (movenet/pytorch_lightning_trainer.py:16,
movenet/trainer.py:29, tests/test_model.py:9 of cosmicBboy/movenet) resolves to the B200-native
implementation in ``movenet_b200`` when this repository is on ``sys.path`` ahead of the reference, so
the trainers run with their import lines unchanged.  Only the hot-path modules exist here
(``wavenet``, ``modules``, ``types``); everything else of the reference package (trainers, config,
dataset, callbacks) is out of scope and is NOT shadowed: put the reference's own files next to these
three, or extend ``movenet.__path__``, to use them."""
import os as _os

# let the reference's own out-of-scope modules (config, dataset, trainers ...) resolve from a checkout named by
# $MOVENET_REFERENCE_DIR without shadowing the three hot-path modules above
_ref = _os.environ.get("MOVENET_REFERENCE_DIR")
if _ref and _os.path.isdir(_os.path.join(_ref, "movenet")):
    __path__.append(_os.path.join(_ref, "movenet"))
