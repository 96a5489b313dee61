"""CPU checker for the NMF hot path -- test infrastructure, never imported by the product package."""
