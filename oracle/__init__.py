"""CPU oracle for the RP-Net hot path — test infrastructure, never imported by the product."""
