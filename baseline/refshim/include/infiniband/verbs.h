/* Interface declarations (subset) of the libibverbs API, for building the unmodified
 * bd-iaas-us/infiniStore reference on an image that has neither rdma-core nor a NIC.
 * Names, enumerators and struct members follow the public verbs API; the implementation behind
 * them here is baseline/refshim/noverbs.c: a NULL provider with one pseudo device that
 * supports only protection domains and memory-region bookkeeping (the reference registers
 * its pool unconditionally at start-up) - queue pairs cannot be created, so the reference's
 * RDMA data path is unavailable (there is no RDMA hardware to run it on); its LOCAL_GPU data
 * path (TCP + CUDA IPC + cudaMemcpyAsync) never touches verbs and runs unmodified.
 */
#ifndef REFSHIM_INFINIBAND_VERBS_H
#define REFSHIM_INFINIBAND_VERBS_H

#include <errno.h>
#include <pthread.h>
#include <stddef.h>
#include <stdint.h>
#include <string.h>
#include <sys/types.h>

#ifdef __cplusplus
extern "C" {
#endif

union ibv_gid {
    uint8_t raw[16];
    struct {
        uint64_t subnet_prefix;
        uint64_t interface_id;
    } global;
};

enum ibv_gid_type { IBV_GID_TYPE_IB_ROCE_V1 = 0, IBV_GID_TYPE_ROCE_V1 = 0, IBV_GID_TYPE_ROCE_V2 = 1 };

enum ibv_mtu { IBV_MTU_256 = 1, IBV_MTU_512 = 2, IBV_MTU_1024 = 3, IBV_MTU_2048 = 4, IBV_MTU_4096 = 5 };
enum ibv_port_state { IBV_PORT_NOP = 0, IBV_PORT_DOWN = 1, IBV_PORT_INIT = 2, IBV_PORT_ARMED = 3, IBV_PORT_ACTIVE = 4 };
enum { IBV_LINK_LAYER_UNSPECIFIED = 0, IBV_LINK_LAYER_INFINIBAND = 1, IBV_LINK_LAYER_ETHERNET = 2 };

struct ibv_context;
struct ibv_device {
    void* ops_[2];
    int node_type;
    int transport_type;
    char name[64];
    char dev_name[64];
    char dev_path[256];
    char ibdev_path[256];
};
struct ibv_context {
    struct ibv_device* device;
    int cmd_fd;
    int async_fd;
    int num_comp_vectors;
    void* abi_compat;
};
struct ibv_pd {
    struct ibv_context* context;
    uint32_t handle;
};
struct ibv_mr {
    struct ibv_context* context;
    struct ibv_pd* pd;
    void* addr;
    size_t length;
    uint32_t handle;
    uint32_t lkey;
    uint32_t rkey;
};
struct ibv_comp_channel {
    struct ibv_context* context;
    int fd;
    int refcnt;
};
struct ibv_cq {
    struct ibv_context* context;
    struct ibv_comp_channel* channel;
    void* cq_context;
    uint32_t handle;
    int cqe;
};
struct ibv_srq;
struct ibv_qp {
    struct ibv_context* context;
    void* qp_context;
    struct ibv_pd* pd;
    struct ibv_cq* send_cq;
    struct ibv_cq* recv_cq;
    struct ibv_srq* srq;
    uint32_t handle;
    uint32_t qp_num;
    int state;
    int qp_type;
};

struct ibv_port_attr {
    enum ibv_port_state state;
    enum ibv_mtu max_mtu;
    enum ibv_mtu active_mtu;
    int gid_tbl_len;
    uint32_t port_cap_flags;
    uint32_t max_msg_sz;
    uint32_t bad_pkey_cntr;
    uint32_t qkey_viol_cntr;
    uint16_t pkey_tbl_len;
    uint16_t lid;
    uint16_t sm_lid;
    uint8_t lmc;
    uint8_t max_vl_num;
    uint8_t sm_sl;
    uint8_t subnet_timeout;
    uint8_t init_type_reply;
    uint8_t active_width;
    uint8_t active_speed;
    uint8_t phys_state;
    uint8_t link_layer;
    uint8_t flags;
    uint16_t port_cap_flags2;
};

enum ibv_access_flags {
    IBV_ACCESS_LOCAL_WRITE = 1,
    IBV_ACCESS_REMOTE_WRITE = (1 << 1),
    IBV_ACCESS_REMOTE_READ = (1 << 2),
    IBV_ACCESS_REMOTE_ATOMIC = (1 << 3),
};
enum ibv_qp_type { IBV_QPT_RC = 2, IBV_QPT_UC, IBV_QPT_UD };
enum ibv_qp_state { IBV_QPS_RESET, IBV_QPS_INIT, IBV_QPS_RTR, IBV_QPS_RTS, IBV_QPS_SQD, IBV_QPS_SQE, IBV_QPS_ERR };
enum ibv_qp_attr_mask {
    IBV_QP_STATE = 1 << 0,
    IBV_QP_CUR_STATE = 1 << 1,
    IBV_QP_EN_SQD_ASYNC_NOTIFY = 1 << 2,
    IBV_QP_ACCESS_FLAGS = 1 << 3,
    IBV_QP_PKEY_INDEX = 1 << 4,
    IBV_QP_PORT = 1 << 5,
    IBV_QP_QKEY = 1 << 6,
    IBV_QP_AV = 1 << 7,
    IBV_QP_PATH_MTU = 1 << 8,
    IBV_QP_TIMEOUT = 1 << 9,
    IBV_QP_RETRY_CNT = 1 << 10,
    IBV_QP_RNR_RETRY = 1 << 11,
    IBV_QP_RQ_PSN = 1 << 12,
    IBV_QP_MAX_QP_RD_ATOMIC = 1 << 13,
    IBV_QP_ALT_PATH = 1 << 14,
    IBV_QP_MIN_RNR_TIMER = 1 << 15,
    IBV_QP_SQ_PSN = 1 << 16,
    IBV_QP_MAX_DEST_RD_ATOMIC = 1 << 17,
    IBV_QP_PATH_MIG_STATE = 1 << 18,
    IBV_QP_CAP = 1 << 19,
    IBV_QP_DEST_QPN = 1 << 20,
};
struct ibv_qp_cap {
    uint32_t max_send_wr;
    uint32_t max_recv_wr;
    uint32_t max_send_sge;
    uint32_t max_recv_sge;
    uint32_t max_inline_data;
};
struct ibv_qp_init_attr {
    void* qp_context;
    struct ibv_cq* send_cq;
    struct ibv_cq* recv_cq;
    struct ibv_srq* srq;
    struct ibv_qp_cap cap;
    enum ibv_qp_type qp_type;
    int sq_sig_all;
};
struct ibv_global_route {
    union ibv_gid dgid;
    uint32_t flow_label;
    uint8_t sgid_index;
    uint8_t hop_limit;
    uint8_t traffic_class;
};
struct ibv_ah_attr {
    struct ibv_global_route grh;
    uint16_t dlid;
    uint8_t sl;
    uint8_t src_path_bits;
    uint8_t static_rate;
    uint8_t is_global;
    uint8_t port_num;
};
struct ibv_qp_attr {
    enum ibv_qp_state qp_state;
    enum ibv_qp_state cur_qp_state;
    enum ibv_mtu path_mtu;
    int path_mig_state;
    uint32_t qkey;
    uint32_t rq_psn;
    uint32_t sq_psn;
    uint32_t dest_qp_num;
    unsigned int qp_access_flags;
    struct ibv_qp_cap cap;
    struct ibv_ah_attr ah_attr;
    struct ibv_ah_attr alt_ah_attr;
    uint16_t pkey_index;
    uint16_t alt_pkey_index;
    uint8_t en_sqd_async_notify;
    uint8_t sq_draining;
    uint8_t max_rd_atomic;
    uint8_t max_dest_rd_atomic;
    uint8_t min_rnr_timer;
    uint8_t port_num;
    uint8_t timeout;
    uint8_t retry_cnt;
    uint8_t rnr_retry;
    uint8_t alt_port_num;
    uint8_t alt_timeout;
    uint32_t rate_limit;
};

enum ibv_wr_opcode {
    IBV_WR_RDMA_WRITE,
    IBV_WR_RDMA_WRITE_WITH_IMM,
    IBV_WR_SEND,
    IBV_WR_SEND_WITH_IMM,
    IBV_WR_RDMA_READ,
};
enum ibv_send_flags {
    IBV_SEND_FENCE = 1 << 0,
    IBV_SEND_SIGNALED = 1 << 1,
    IBV_SEND_SOLICITED = 1 << 2,
    IBV_SEND_INLINE = 1 << 3,
};
struct ibv_sge {
    uint64_t addr;
    uint32_t length;
    uint32_t lkey;
};
struct ibv_send_wr {
    uint64_t wr_id;
    struct ibv_send_wr* next;
    struct ibv_sge* sg_list;
    int num_sge;
    enum ibv_wr_opcode opcode;
    unsigned int send_flags;
    union {
        uint32_t imm_data;
        uint32_t invalidate_rkey;
    };
    union {
        struct {
            uint64_t remote_addr;
            uint32_t rkey;
        } rdma;
        struct {
            uint64_t remote_addr;
            uint64_t compare_add;
            uint64_t swap;
            uint32_t rkey;
        } atomic;
    } wr;
};
struct ibv_recv_wr {
    uint64_t wr_id;
    struct ibv_recv_wr* next;
    struct ibv_sge* sg_list;
    int num_sge;
};
enum ibv_wc_status { IBV_WC_SUCCESS = 0, IBV_WC_LOC_LEN_ERR, IBV_WC_GENERAL_ERR = 21 };
enum ibv_wc_opcode {
    IBV_WC_SEND,
    IBV_WC_RDMA_WRITE,
    IBV_WC_RDMA_READ,
    IBV_WC_COMP_SWAP,
    IBV_WC_FETCH_ADD,
    IBV_WC_BIND_MW,
    IBV_WC_RECV = 1 << 7,
    IBV_WC_RECV_RDMA_WITH_IMM,
};
struct ibv_wc {
    uint64_t wr_id;
    enum ibv_wc_status status;
    enum ibv_wc_opcode opcode;
    uint32_t vendor_err;
    uint32_t byte_len;
    union {
        uint32_t imm_data;
        uint32_t invalidated_rkey;
    };
    uint32_t qp_num;
    uint32_t src_qp;
    unsigned int wc_flags;
    uint16_t pkey_index;
    uint16_t slid;
    uint8_t sl;
    uint8_t dlid_path_bits;
};

struct ibv_device** ibv_get_device_list(int* num_devices);
void ibv_free_device_list(struct ibv_device** list);
const char* ibv_get_device_name(struct ibv_device* device);
struct ibv_context* ibv_open_device(struct ibv_device* device);
int ibv_close_device(struct ibv_context* context);
int ibv_query_port(struct ibv_context* context, uint8_t port_num, struct ibv_port_attr* port_attr);
int ibv_query_gid(struct ibv_context* context, uint8_t port_num, int index, union ibv_gid* gid);
struct ibv_pd* ibv_alloc_pd(struct ibv_context* context);
int ibv_dealloc_pd(struct ibv_pd* pd);
struct ibv_mr* ibv_reg_mr(struct ibv_pd* pd, void* addr, size_t length, int access);
int ibv_dereg_mr(struct ibv_mr* mr);
struct ibv_comp_channel* ibv_create_comp_channel(struct ibv_context* context);
int ibv_destroy_comp_channel(struct ibv_comp_channel* channel);
struct ibv_cq* ibv_create_cq(struct ibv_context* context, int cqe, void* cq_context,
                             struct ibv_comp_channel* channel, int comp_vector);
int ibv_destroy_cq(struct ibv_cq* cq);
int ibv_get_cq_event(struct ibv_comp_channel* channel, struct ibv_cq** cq, void** cq_context);
void ibv_ack_cq_events(struct ibv_cq* cq, unsigned int nevents);
int ibv_req_notify_cq(struct ibv_cq* cq, int solicited_only);
int ibv_poll_cq(struct ibv_cq* cq, int num_entries, struct ibv_wc* wc);
struct ibv_qp* ibv_create_qp(struct ibv_pd* pd, struct ibv_qp_init_attr* qp_init_attr);
int ibv_modify_qp(struct ibv_qp* qp, struct ibv_qp_attr* attr, int attr_mask);
int ibv_destroy_qp(struct ibv_qp* qp);
int ibv_post_send(struct ibv_qp* qp, struct ibv_send_wr* wr, struct ibv_send_wr** bad_wr);
int ibv_post_recv(struct ibv_qp* qp, struct ibv_recv_wr* wr, struct ibv_recv_wr** bad_wr);
const char* ibv_wc_status_str(enum ibv_wc_status status);

#ifdef __cplusplus
}
#endif
#endif
