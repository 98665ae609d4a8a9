"""Import shim (test infrastructure): the reference imports three helpers from timm, which is absent here."""
