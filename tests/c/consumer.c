/* A plain-C consumer of the C-ABI (include/hd_b200.h): proves the header is C-clean and that the network-level entries can be
 * driven without Python.  Built and run by tests/test_abi.py::test_plain_c_consumer (no GPU needed: the weight callback reports a
 * missing variable, which hd_resnet50_create must turn into HD_ERR_INVALID naming it before it touches the device). */
#include <dlfcn.h>
#include <stdio.h>
#include <string.h>

#include "hd_b200.h"

static int asked = 0;
static char first_name[256];

static const float *no_weights(void *user, const char *tf_name, long long *numel) {
  (void)user; (void)numel;
  if (!asked++) { strncpy(first_name, tf_name, sizeof(first_name) - 1); }
  return NULL;
}

typedef int (*version_fn)(void);
typedef const char *(*str_fn)(void);
typedef const char *(*status_fn)(int);
typedef int (*create_fn)(hd_weight_fn, void *, int, int, hd_net **);
typedef void (*destroy_fn)(hd_net *);
typedef int (*geom_fn)(int, int, const double *, int, int *, int *, int *);

int main(int argc, char **argv) {
  if (argc < 2) return 2;
  void *h = dlopen(argv[1], RTLD_NOW);
  if (!h) { fprintf(stderr, "dlopen: %s\n", dlerror()); return 3; }
  version_fn version = (version_fn)dlsym(h, "hd_version");
  str_fn last_error = (str_fn)dlsym(h, "hd_last_error");
  status_fn status_string = (status_fn)dlsym(h, "hd_status_string");
  create_fn create = (create_fn)dlsym(h, "hd_resnet50_create");
  destroy_fn destroy = (destroy_fn)dlsym(h, "hd_net_destroy");
  geom_fn crop_geometry = (geom_fn)dlsym(h, "hd_crop_geometry");
  if (!version || !last_error || !status_string || !create || !destroy || !crop_geometry) return 4;
  {
    /* run_video.py:69-100 for a 240x320 frame, bbox centre (40.3, 200.7), scale 0.62 (case 1 of tests/golden/preproc_cases.py) */
    const double bbox[3] = {40.3, 200.7, 0.62};
    int geom[4], center[2], start[2];
    int grc = crop_geometry(240, 320, bbox, 224, geom, center, start);
    printf("geom rc=%d %d %d %d %d center=%d,%d start=%d,%d\n", grc, geom[0], geom[1], geom[2], geom[3], center[0], center[1], start[0], start[1]);
    if (grc != HD_OK) return 5;
  }
  hd_net *net = (hd_net *)0x1;
  hd_conv_desc d;
  memset(&d, 0, sizeof(d));
  d.out_subsample = 2;                       /* the struct is usable from C as declared */
  int rc = create(no_weights, NULL, 4, 224, &net);
  printf("version=%d rc=%d status=%s net=%p asked=%d first=%s error=%s\n", version(), rc, status_string(rc), (void *)net, asked,
         first_name, last_error());
  destroy(net);                              /* NULL: must be a no-op */
  return (rc == HD_ERR_INVALID && net == NULL && strstr(last_error(), first_name) != NULL) ? 0 : 1;
}
