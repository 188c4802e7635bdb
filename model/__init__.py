"""Drop-in replacement of the reference `model` package for the dense-regression hot path: `from model.main_model import
mainModel` (reference main.py:19) resolves here, with the same constructor, attribute tree, state_dict keys and forward
signature, but the work is done by the sm_100a kernels of drn_b200/libdrn_sm100.so."""
