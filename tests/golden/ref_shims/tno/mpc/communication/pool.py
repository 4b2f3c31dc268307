from tno.mpc.communication import Pool  # noqa: F401
