"""Importable alias for the product package.

The package directory is named after the reference repository
(``multimodal-prediction-and-cervical-lesion-slice-segmentation-based-on-deep-learning_b200``),
which is not a valid Python identifier; this shim extends ``__path__`` so that
``import cervix_b200.ops`` etc. resolve to the modules in that directory.
"""
import os as _os

PACKAGE_DIR = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))),
                            "multimodal-prediction-and-cervical-lesion-slice-segmentation-based-on-deep-learning_b200")
__path__.append(PACKAGE_DIR)
