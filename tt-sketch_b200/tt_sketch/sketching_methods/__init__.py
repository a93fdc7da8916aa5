r"""Omega / Psi contractions per tensor type.

For each input type there is a pair ``sketch_omega_<type>`` / ``sketch_psi_<type>`` with the
reference's NumPy-in / NumPy-out signature (the plug-in point registered in
``sketch_dispatch.OMEGA_METHODS`` / ``PSI_METHODS``) and a device-resident pair
``omega_<type>_device`` / ``psi_<type>_device`` that `general_sketch` uses; both run the same
CUDA kernels.
"""
