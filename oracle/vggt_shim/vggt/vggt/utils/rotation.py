from oracle.functional import quat_to_mat, mat_to_quat  # noqa: F401
