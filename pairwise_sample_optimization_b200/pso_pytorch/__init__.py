"""Mirror of the reference's ``pso_pytorch`` package layout (human_preference_tuning/pso_pytorch):
same module and function names for the hot-path entry points, backed by the sm_100a kernels."""
