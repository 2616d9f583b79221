"""Overlay of the reference's modules/dynamic_modules/budget.py."""
from dynamicvectorquantization_b200.nn.router import (BudgetConstraint_NormedSeperateRatioMSE_TripleGrain,  # noqa: F401
                                                      BudgetConstraint_RatioMSE_DualGrain)
