/* NULL libibverbs provider (see include/infiniband/verbs.h in this directory).
 * One pseudo device so that the reference server's unconditional RDMA context set-up
 * (device list, port query, protection domain, registration of its host pool) succeeds on a
 * box without RDMA hardware.  Nothing here moves data: queue pairs cannot be created and
 * posting work fails with ENODEV, i.e. the reference's RDMA path reports "unavailable" and
 * only its LOCAL_GPU path - which uses no verbs at all - can run. */
#include <errno.h>
#include <stdlib.h>
#include <string.h>
#include <sys/eventfd.h>
#include <unistd.h>

#include "infiniband/verbs.h"

static struct ibv_device g_dev;
static struct ibv_device* g_list[2];

struct ibv_device** ibv_get_device_list(int* num_devices) {
    memset(&g_dev, 0, sizeof(g_dev));
    strcpy(g_dev.name, "nonic0");
    strcpy(g_dev.dev_name, "nonic0");
    strcpy(g_dev.ibdev_path, "/nonexistent/nonic0");
    g_list[0] = &g_dev;
    g_list[1] = NULL;
    if (num_devices) *num_devices = 1;
    return g_list;
}
void ibv_free_device_list(struct ibv_device** list) { (void)list; }
const char* ibv_get_device_name(struct ibv_device* device) { return device ? device->name : NULL; }
struct ibv_context* ibv_open_device(struct ibv_device* device) {
    struct ibv_context* c = calloc(1, sizeof(*c));
    if (c) c->device = device;
    return c;
}
int ibv_close_device(struct ibv_context* context) {
    free(context);
    return 0;
}
int ibv_query_port(struct ibv_context* context, uint8_t port_num, struct ibv_port_attr* a) {
    (void)context;
    (void)port_num;
    memset(a, 0, sizeof(*a));
    a->state = IBV_PORT_ACTIVE;
    a->max_mtu = a->active_mtu = IBV_MTU_4096;
    a->lid = 1;
    a->link_layer = IBV_LINK_LAYER_INFINIBAND; /* no GID lookup needed for "IB" */
    return 0;
}
int ibv_query_gid(struct ibv_context* context, uint8_t port_num, int index, union ibv_gid* gid) {
    (void)context;
    (void)port_num;
    (void)index;
    memset(gid, 0, sizeof(*gid));
    return 0;
}
struct ibv_pd* ibv_alloc_pd(struct ibv_context* context) {
    struct ibv_pd* pd = calloc(1, sizeof(*pd));
    if (pd) pd->context = context;
    return pd;
}
int ibv_dealloc_pd(struct ibv_pd* pd) {
    free(pd);
    return 0;
}
struct ibv_mr* ibv_reg_mr(struct ibv_pd* pd, void* addr, size_t length, int access) {
    static uint32_t next_key = 1;
    (void)access;
    struct ibv_mr* mr = calloc(1, sizeof(*mr));
    if (!mr) return NULL;
    mr->context = pd ? pd->context : NULL;
    mr->pd = pd;
    mr->addr = addr;
    mr->length = length;
    mr->lkey = mr->rkey = next_key++;
    return mr;
}
int ibv_dereg_mr(struct ibv_mr* mr) {
    free(mr);
    return 0;
}
struct ibv_comp_channel* ibv_create_comp_channel(struct ibv_context* context) {
    struct ibv_comp_channel* ch = calloc(1, sizeof(*ch));
    if (!ch) return NULL;
    ch->context = context;
    ch->fd = eventfd(0, EFD_NONBLOCK | EFD_CLOEXEC); /* never becomes readable */
    return ch;
}
int ibv_destroy_comp_channel(struct ibv_comp_channel* ch) {
    if (ch) {
        if (ch->fd >= 0) close(ch->fd);
        free(ch);
    }
    return 0;
}
struct ibv_cq* ibv_create_cq(struct ibv_context* context, int cqe, void* cq_context,
                             struct ibv_comp_channel* channel, int comp_vector) {
    (void)comp_vector;
    struct ibv_cq* cq = calloc(1, sizeof(*cq));
    if (!cq) return NULL;
    cq->context = context;
    cq->channel = channel;
    cq->cq_context = cq_context;
    cq->cqe = cqe;
    return cq;
}
int ibv_destroy_cq(struct ibv_cq* cq) {
    free(cq);
    return 0;
}
int ibv_get_cq_event(struct ibv_comp_channel* channel, struct ibv_cq** cq, void** cq_context) {
    (void)channel;
    (void)cq;
    (void)cq_context;
    errno = EAGAIN;
    return -1;
}
void ibv_ack_cq_events(struct ibv_cq* cq, unsigned int nevents) {
    (void)cq;
    (void)nevents;
}
int ibv_req_notify_cq(struct ibv_cq* cq, int solicited_only) {
    (void)cq;
    (void)solicited_only;
    return 0;
}
int ibv_poll_cq(struct ibv_cq* cq, int num_entries, struct ibv_wc* wc) {
    (void)cq;
    (void)num_entries;
    (void)wc;
    return 0; /* no completions, ever */
}
struct ibv_qp* ibv_create_qp(struct ibv_pd* pd, struct ibv_qp_init_attr* a) {
    (void)pd;
    (void)a;
    errno = ENODEV; /* there is no RDMA hardware behind this provider */
    return NULL;
}
int ibv_modify_qp(struct ibv_qp* qp, struct ibv_qp_attr* attr, int attr_mask) {
    (void)qp;
    (void)attr;
    (void)attr_mask;
    return ENODEV;
}
int ibv_destroy_qp(struct ibv_qp* qp) {
    (void)qp;
    return 0;
}
int ibv_post_send(struct ibv_qp* qp, struct ibv_send_wr* wr, struct ibv_send_wr** bad_wr) {
    (void)qp;
    if (bad_wr) *bad_wr = wr;
    return ENODEV;
}
int ibv_post_recv(struct ibv_qp* qp, struct ibv_recv_wr* wr, struct ibv_recv_wr** bad_wr) {
    (void)qp;
    if (bad_wr) *bad_wr = wr;
    return ENODEV;
}
const char* ibv_wc_status_str(enum ibv_wc_status status) {
    return status == IBV_WC_SUCCESS ? "success" : "error (null verbs provider)";
}
