"""Drop-in front-ends with the argv / stdin contracts of the reference's closed binaries
(SURVEY.md Appendix A): refine3d, reconstruct3d, local_merge3d, merge3d, csp."""
