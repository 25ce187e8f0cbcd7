from .pwclo_net import PWCLONet, PWCLONET  # noqa: F401
