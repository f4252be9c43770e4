from .auto_policy import get_autopolicy, import_policy, register_policy
from .base_policy import ModulePolicyDescription, Policy, SubModuleReplacementDescription
from . import zoo  # noqa: F401,E402  (generates `policies.<family>` for the zoo's text families)

__all__ = ["get_autopolicy", "import_policy", "register_policy", "ModulePolicyDescription", "Policy",
           "SubModuleReplacementDescription"]
