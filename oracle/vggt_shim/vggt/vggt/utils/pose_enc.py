from oracle.functional import extri_intri_to_pose_encoding, pose_encoding_to_extri_intri  # noqa: F401
