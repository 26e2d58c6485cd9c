"""TEST INFRASTRUCTURE ONLY -- stand-in for the un-vendored FrEIA v0.2 dependency.

The reference pins `FrEIA @ c5fe1af0de8ce9122b5b61924ad75a19b9dc2473`
(/root/reference/requirements.txt:2).  That package is not on disk and there is
no network, so this stand-in restates the published behaviour of the handful of
FrEIA classes the reference touches.  It exists so that
`oracle/make_golden.py` can import the reference's *unmodified* files in this
container.  "parity unpinned" for everything that lives in here: no FrEIA
source or FrEIA test vector was available to check it against.
"""
