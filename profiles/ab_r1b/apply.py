import sys
sys.path.insert(0, '/tmp/var')
import patches
d = sys.argv[1]
for name in sys.argv[2:]:
    getattr(patches, name)(d)
    print('applied', name)
