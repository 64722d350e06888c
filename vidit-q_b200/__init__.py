"""viditq_b200 — B200-native low-bit path behind ViDiT-Q's qdiff QuantLayer operator (import as `viditq_b200`)."""
