from .auto_policy import get_autopolicy, import_policy, register_policy
from .base_policy import ModulePolicyDescription, Policy, SubModuleReplacementDescription

__all__ = ["get_autopolicy", "import_policy", "register_policy", "ModulePolicyDescription", "Policy",
           "SubModuleReplacementDescription"]
