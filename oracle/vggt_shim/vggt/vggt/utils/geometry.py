from oracle.functional import closed_form_inverse_se3  # noqa: F401
