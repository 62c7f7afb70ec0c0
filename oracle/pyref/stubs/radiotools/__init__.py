# Import-time stand-in for the third-party `radiotools` package (absent in this image).
# Only what the reference's import chain touches is provided; none of it is on the ray-tracing hot path.
