"""TEST INFRASTRUCTURE ONLY -- a stand-in for the parts of astropy the reference touches on the
hot path (``astropy.units`` arithmetic between Hz/MHz/s/cycle/rad/pc/cm and a scalar
``astropy.time.Time``).  astropy is not installable in the build image; this stub exists so that
the reference's OWN source files under /root/reference can be executed here, unmodified, to
produce golden vectors (oracle/make_ref_golden.py).  It is never imported by the product."""

__version__ = "0.0-stub"
