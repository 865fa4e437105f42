"""bluetangle.jl_b200 -- B200-native backend for BlueTangle.jl's apply/noise/measure/sample/expect path.

``csrc/`` holds the hand-written sm_100a kernels and the C ABI (include/bluetangle_cuda.h),
``host.py`` the host-side mirror of the reference's front end for device-resident states.
The directory name contains a dot, so load it with ``__graft_entry__.load_package()``.
"""
from . import _lib
from .host import *  # noqa: F401,F403
from .host import _born_measure, _reset_Z, _final_measurement, _sample_to_expectation  # noqa: F401
from .gates import gate, gates, noise_model, is_valid_quantum_channel  # noqa: F401
from .qasm import from_qasm  # noqa: F401,E402
from .vqa import (PauliSum, hamiltonian, VOp, AnsatzOptions, variational_apply, loss_and_grad_paramshift, loss_and_grad, shift_rule, VQA, VQE,  # noqa: F401,E402
                  EfficientSU2, generate_ansatz_circuit, _variational_circuit_from_string)
