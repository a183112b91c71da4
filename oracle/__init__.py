"""oracle/ -- TEST INFRASTRUCTURE ONLY (parity checkers). Never imported by the product package."""
