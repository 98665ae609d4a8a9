"""Host-side mirrors of the reference's ``models`` package for the denoising hot path."""
