"""slicetypecheck: run the static Func-call analyzer over python files
(reference ``cmd/slicetypecheck/main.go:14-17``)."""
import sys

from ..analysis.typecheck import main

if __name__ == "__main__":
    sys.exit(main())
