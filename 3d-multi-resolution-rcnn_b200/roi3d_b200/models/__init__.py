"""Host mirrors of the reference models that sit on the 3D RoI hot path (extractor, RPN proposal path, mask paste)."""
